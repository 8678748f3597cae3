/* ddpb.h -- C ABI of the B200-native batched DDP hot path.
 *
 * The reference (s-elkazdadi/ddp-pinocchio) is a header-only C++ template library without any
 * FFI layer; its boundary for this path is three duck-typed concepts.
 * Each entry point below names the reference interface it replaces.  A batch is B independent
 * problems that share one model, horizon T, time step, cost weight and constraint structure
 * and differ in their trajectories, targets and multipliers.
 *
 * Conventions
 *  - all host arrays are double, problem-major: problem b's block follows problem b-1's;
 *  - inside a problem block every sequence uses the reference's packed mat_seq layout
 *    (include/ddp/detail/mat_seq.hpp:10-128): knot t at offset sum_{t'<t} rows(t')*cols(t'),
 *    matrices column-major, tensors T(k,i,j) at k + i*O + j*O*L (detail/tensor.hpp:141-147);
 *  - every function returns DDPB_OK or an error code, never throws, and fails with
 *    DDPB_ERR_CUDA when no usable sm_100 device is present (there is no CPU fallback);
 *  - a handle may be used from one host thread at a time; different handles are independent.
 */
#ifndef DDPB_H_
#define DDPB_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ddpb_model ddpb_model;
typedef struct ddpb_batch ddpb_batch;

enum ddpb_status { DDPB_OK = 0, DDPB_ERR_INVALID = 1, DDPB_ERR_UNSUPPORTED = 2, DDPB_ERR_CUDA = 3, DDPB_ERR_NOMEM = 4 };
enum ddpb_model_kind {
  DDPB_MODEL_PENDULUM = 0,       /* pendulum_model_t, include/ddp/pendulum_model.hpp:12-132 */
  DDPB_MODEL_REVOLUTE_CHAIN = 1  /* pinocchio::model_t restricted to a serial chain of revolute / prismatic joints */
};
enum ddpb_joint_type {           /* what pinocchio::urdf::buildModel (pinocchio_model.ipp:111) makes of URDF joints */
  DDPB_JOINT_REVOLUTE = 0,       /* rotation about `axis`: motion subspace [axis; 0] in the joint frame */
  DDPB_JOINT_PRISMATIC = 1       /* translation along `axis`: motion subspace [0; axis]; nq = nv = 1 as well */
};
enum ddpb_eq_kind {
  DDPB_EQ_NONE = 0,
  DDPB_EQ_CONFIG = 1,  /* advance(advance(config_constraint_t)),  problem.hpp:744-864 + 527-624 */
  DDPB_EQ_SPATIAL = 2  /* advance(advance(spatial_constraint_t)), problem.hpp:631-742 + 527-624 */
};
enum ddpb_hess_mode {
  DDPB_HESS_FD_JACOBIANS = 0, /* dynamics_t::second_order_finite_diff = false, problem.hpp:67-150  */
  DDPB_HESS_FD_VALUES = 1     /* dynamics_t::second_order_finite_diff = true,  problem.hpp:152-298 */
};
enum ddpb_frame_jac_mode {
  DDPB_FJ_WORLD = 0,              /* what the reference uses: getFrameJacobian(..., WORLD), pinocchio_model.ipp:460 */
  DDPB_FJ_LOCAL_WORLD_ALIGNED = 1 /* true d p_frame / d q (opt-in; changes results) */
};

/* Replaces pinocchio::urdf::buildModel + ModelTpl (pinocchio_model.ipp:98-152) / pendulum_model_t's constructor. */
typedef struct ddpb_model_desc {
  int kind;                      /* ddpb_model_kind */
  int n_joints;                  /* nq = nv = n_joints (1..7); ignored for the pendulum */
  const int* parent;             /* [n] parent joint, -1 = fixed base; must be i-1 (serial chain) */
  const double* placement_rot;   /* [n][9] row-major rotation joint frame -> parent coordinates */
  const double* placement_trans; /* [n][3] */
  const double* axis;            /* [n][3] unit joint axis in the joint frame */
  const double* mass;            /* [n] */
  const double* com;             /* [n][3] centre of mass in the body frame */
  const double* inertia;         /* [n][6] xx,xy,xz,yy,yz,zz about the centre of mass */
  double gravity[3];             /* (0,0,-9.81) */
  int n_frames;                  /* operational frames (<= 16) */
  const int* frame_parent;       /* [f] supporting joint */
  const double* frame_rot;       /* [f][9] (unused by the path: only translations enter) */
  const double* frame_trans;     /* [f][3] */
  double pend_mass, pend_length; /* pendulum_model_t{mass, length} */
  const int* joint_type;         /* [n] ddpb_joint_type; NULL = all revolute (last member: a zero-initialised descriptor
                                  * written before this field existed keeps its meaning) */
} ddpb_model_desc;

/* dynamics_t{model, dt, second_order_finite_diff} (problem.hpp:522-524), problem_t{begin,end,c,...}
 * (problem.hpp:1145-1147) and the constraint's structure. */
typedef struct ddpb_params {
  double dt;
  double c;           /* l(t,x,u) = c/2 |u|^2, lf = 0 (problem.hpp:932-942) */
  int hess_mode;      /* ddpb_hess_mode */
  int eq_kind;        /* ddpb_eq_kind */
  int frame_id;       /* DDPB_EQ_SPATIAL: constrained frame (test/pinocchio_spatial_eq_ddp.cpp:81) */
  int frame_jac_mode; /* ddpb_frame_jac_mode */
  /* 0 = the reference's behaviour: after a failed Cholesky the backward pass doubles mu and restarts, without
   * bound (ddp_bwd.ipp:105-132), and a problem whose mu overflowed keeps producing NaN trajectories.
   * N > 0 = guarded mode: a problem that needs more than N restarts in ONE backward pass, or whose mu / merit
   * stop being finite, gets status 2 (diverged) and is left untouched from then on. */
  int max_backward_restarts;
} ddpb_params;

/* solver_parameters_t (ddp.hpp:42-50) */
typedef struct ddpb_solver_params {
  int max_iterations;
  double optimality_stopping_threshold;
  double mu, reg, w, n;
} ddpb_solver_params;

/* the 17 sequences of derivative_storage_t plus lfx, lfxx (ddp.hpp:84-105); each pointer is
 * [B][packed sequence] on the host, NULL = skip */
typedef struct ddpb_derivs {
  double *lfx, *lfxx, *lx, *lu, *lxx, *lux, *luu, *f_val, *fx, *fu, *fxx, *fux, *fuu, *eq_val, *eq_x, *eq_u, *eq_xx, *eq_ux, *eq_uu;
} ddpb_derivs;

/* per-problem scalars, field-major [DDPB_NSCALARS][B] */
enum ddpb_scalar_field {
  DDPB_S_MU = 0, DDPB_S_REG, DDPB_S_W, DDPB_S_N, DDPB_S_STEP, DDPB_S_OPT_OBJ, DDPB_S_OPT_CONSTR,
  DDPB_S_COST,       /* sum_t l(t,x,u) of the current trajectory */
  DDPB_S_MERIT,      /* sum_t cost_seq_aug of the accepted trajectory (ddp.hpp:699-735) */
  DDPB_S_ITERATIONS, /* DDP iterations completed */
  DDPB_S_STATUS,     /* 0 running / max iterations, 1 optimum attained, 2 diverged (guarded mode only) */
  DDPB_S_ROLLOUTS,   /* line-search rollouts (cumulative) */
  DDPB_S_RESTARTS,   /* backward-pass restarts after a failed Cholesky (cumulative) */
  DDPB_S_MULT_UPDATES,
  DDPB_NSCALARS = 16
};

/* ---- lifetime -------------------------------------------------------------------------- */
int ddpb_model_create(const ddpb_model_desc* desc, ddpb_model** out);
/* model_t{urdf_path} -> pinocchio::urdf::buildModel(path, model) with a fixed base (pinocchio_model.ipp:98-113, :155): the same
 * from URDF TEXT (the caller reads the file).  Revolute / continuous / prismatic joints in a serial chain; fixed joints are merged into
 * the body of their supporting joint and their child links become operational frames (every link is a BODY frame, in
 * traversal order); anything else returns DDPB_ERR_UNSUPPORTED, malformed text DDPB_ERR_INVALID. */
int ddpb_model_create_from_urdf(const char* urdf_xml, unsigned long long length, ddpb_model** out);
/* what a model holds (adapters, tests): dims2 = {joints, frames}; any output array may be NULL.  Layouts as in ddpb_model_desc. */
int ddpb_model_dims(const ddpb_model* m, int* dims2);
int ddpb_model_get(const ddpb_model* m, double* placement_rot, double* placement_trans, double* axis, double* mass, double* com,
                   double* inertia, int* frame_parent, double* frame_trans);
int ddpb_model_joint_types(const ddpb_model* m, int* joint_type /* [joints] ddpb_joint_type */);
const char* ddpb_model_frame_name(const ddpb_model* m, int frame); /* "" when the model was built from arrays */
void ddpb_model_destroy(ddpb_model* m);

/* ---- the Model concept's own functions, batched ------------------------------------------------
 * The reference's dynamics and constraints call these once per knot on one shared pinocchio::Data; here every call
 * evaluates n independent samples on the device (one thread per sample, csrc/model_eval.cuh) from / into HOST arrays,
 * sample-major: q, v, tau, a [n][nq]; matrices [n][nv * nv] column-major per sample like the reference's Eigen maps.
 * The solver never uses them (compute_derivatives evaluates whole stencils per knot); they are the reference-facing
 * single-function interface and the direct check of the rigid-body layer.  Pendulum models: the two dynamics functions only. */
/* model_t::dynamics_aba(out_acceleration, q, v, tau) -> pinocchio::aba (pinocchio_model.hpp:147-152, pinocchio_model.ipp:336-357);
 * pendulum_model_t::dynamics_aba (pendulum_model.hpp:105-114) */
int ddpb_model_dynamics_aba(const ddpb_model* m, int device, int n, const double* q, const double* v, const double* tau, double* a);
/* model_t::d_dynamics_aba(out_dq, out_dv, out_dtau, q, v, tau) -> pinocchio::computeABADerivatives (pinocchio_model.hpp:154-161,
 * pinocchio_model.ipp:359-400); pendulum_model.hpp:116-130.  Any output may be NULL (not all three). */
int ddpb_model_d_dynamics_aba(const ddpb_model* m, int device, int n, const double* q, const double* v, const double* tau,
                              double* a_dq, double* a_dv, double* a_dtau);
/* model_t::inverse_dynamics_rnea(out_tau, q, v, a) -> pinocchio::rnea (pinocchio_model.hpp:180-185) */
int ddpb_model_inverse_dynamics_rnea(const ddpb_model* m, int device, int n, const double* q, const double* v, const double* a, double* tau);
/* model_t::frame_coordinates(i, q) -> framesForwardKinematics + oMf[i].translation() (pinocchio_model.hpp:163, pinocchio_model.ipp:417-432);
 * p [n][3] */
int ddpb_model_frame_coordinates(const ddpb_model* m, int device, int n, int frame, const double* q, double* p);
/* model_t::d_frame_coordinates(out[3 x nv], i, q) -> top three rows of getFrameJacobian(..., WORLD) (pinocchio_model.hpp:164,
 * pinocchio_model.ipp:434-462); frame_jac_mode: the frame_jac_mode field of the batch parameters, DDPB_FJ_WORLD reproduces the reference; J [n][3 * nv] column-major */
int ddpb_model_d_frame_coordinates(const ddpb_model* m, int device, int n, int frame, int frame_jac_mode, const double* q, double* J);
/* ne_per_knot[t] (t in [0,T)) = equality rows at knot t, i.e. rows of the doubly advanced
 * constraint that acts on x_{t+2} (indexing::shift_time_idx_t, indexer.hpp:291-318); the same for
 * every problem.  Replaces ddp_solver_t's constructor + uninit_derivative_storage (ddp.hpp:430-514,864-867). */
int ddpb_batch_create(const ddpb_model* m, int B, int T, const int* ne_per_knot, const ddpb_params* params, int device, ddpb_batch** out);
void ddpb_batch_destroy(ddpb_batch* b);
int ddpb_batch_set_stream(ddpb_batch* b, void* cuda_stream); /* cudaStream_t; default: a private non-blocking stream */
int ddpb_batch_sync(ddpb_batch* b);
const char* ddpb_last_error(void);
/* pinned host memory for the transfers */
int ddpb_host_alloc(void** ptr, unsigned long long bytes);
int ddpb_host_free(void* ptr);

/* ---- data in / out (trajectory_t, affine_vector_function_seq_t) ---------------------------- */
int ddpb_set_trajectory(ddpb_batch* b, const double* xs /*[B][(T+1) nx]*/, const double* us /*[B][T nu]*/);
int ddpb_get_trajectory(ddpb_batch* b, double* xs, double* us);
int ddpb_get_new_trajectory(ddpb_batch* b, double* xs, double* us); /* storage written by ddpb_forward_pass */
/* ddp_solver_t::make_trajectory (ddp.hpp:392-415): roll controls out from x0; us == NULL means zero controls */
int ddpb_make_trajectory(ddpb_batch* b, const double* x0 /*[B][nx]*/, const double* us /*[B][T nu] or NULL*/);
int ddpb_set_targets(ddpb_batch* b, const double* targets /*[B][sum ne]*/);
/* multipliers (affine in x): val == NULL -> 0; jac == NULL -> 0; origin := current trajectory (ddp.hpp:759-764) */
int ddpb_set_multipliers(ddpb_batch* b, const double* val /*[B][sum ne]*/, const double* jac /*[B][sum ne*ndx]*/);
int ddpb_get_multipliers(ddpb_batch* b, double* origin /*[B][T nx]*/, double* val, double* jac);
int ddpb_set_feedback(ddpb_batch* b, const double* k /*[B][T nu]*/, const double* K /*[B][T nu*ndx]*/); /* origin := current trajectory */
int ddpb_get_feedback(ddpb_batch* b, double* origin /*[B][T nx]*/, double* k, double* K);
int ddpb_set_derivatives(ddpb_batch* b, const ddpb_derivs* in);
int ddpb_get_derivatives(ddpb_batch* b, const ddpb_derivs* out);
int ddpb_set_scalar(ddpb_batch* b, int field, const double* values /*[B]*/);
int ddpb_get_scalars(ddpb_batch* b, double* out /*[DDPB_NSCALARS][B]*/);
/* device pointer of the scalar block ([DDPB_NSCALARS][B] doubles), for NCCL gathers by the host layer */
int ddpb_scalars_device_ptr(ddpb_batch* b, void** ptr);

/* ---- the hot path ------------------------------------------------------------------------ */
/* problem_t::compute_derivatives (problem.hpp:956-997) for every knot of every problem */
int ddpb_compute_derivatives(ddpb_batch* b);
/* ddp_solver_t::backward_pass (ddp_bwd.ipp:9-155); mu / reg are per-problem, in-out, on the host */
int ddpb_backward_pass(ddpb_batch* b, double* mu /*[B]*/, double* reg /*[B]*/);
/* ddp_solver_t::forward_pass with line search (ddp_fwd.ipp:10-67, cost_seq_aug ddp.hpp:699-735);
 * uses DDPB_S_MU of each problem; writes the new-trajectory storage */
int ddpb_forward_pass(ddpb_batch* b, double* step_out /*[B]*/);
/* cost_seq_aug of the current trajectory, [B][T+1] */
int ddpb_cost_seq_aug(ddpb_batch* b, double* costs);
/* optimality_obj / optimality_constr (ddp.hpp:516-627) of the current trajectory and derivatives */
int ddpb_optimality(ddpb_batch* b, double* opt_obj /*[B]*/, double* opt_constr /*[B]*/);
/* ddp_solver_t::solve<primal_dual_affine_multipliers> (ddp.hpp:744-842) on the whole batch, starting
 * from the current trajectory and multipliers; everything stays on the device.  scalars_out may be NULL. */
int ddpb_solve(ddpb_batch* b, const ddpb_solver_params* p, double* scalars_out /*[DDPB_NSCALARS][B]*/);
/* same, without the final synchronisation / read-back (for stream-ordered use and CUDA-event timing) */
int ddpb_solve_async(ddpb_batch* b, const ddpb_solver_params* p);

/* ---- device-resident restart (bench / repeated solves without host traffic) ------------------ */
/* keep a device-side copy of the current trajectory and multipliers / copy it back */
int ddpb_snapshot_save(ddpb_batch* b);
int ddpb_snapshot_restore(ddpb_batch* b);
/* FP64 FMA micro-benchmark on `device` (2 flops per DFMA), TFLOP/s: the denominator of the FP64 roofline */
int ddpb_measure_fp64_peak(int device, double* tflops);

/* ---- introspection ----------------------------------------------------------------------- */
int ddpb_batch_dims(const ddpb_batch* b, int* out /* nq, nx, ndx, nu, ne_total, T, B */);
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
unsigned long long ddpb_launch_count(const ddpb_batch* b);
/* device time of the last ddpb_solve / hot-path call per phase, ms: linearize, backward, forward, other */
int ddpb_phase_times(ddpb_batch* b, float* ms4);
int ddpb_enable_phase_timing(ddpb_batch* b, int on);
/* ddpb_solve runs the batch as `n` contiguous sub-batches (1..32), each on its own pair of CUDA streams, so that the
 * latency-bound sweeps of one sub-batch overlap the linearization of another; results do not depend on `n` (problems
 * are independent -- the reference, ddp.hpp:744-842, solves one problem at a time).  n = 1: one stream, kernels of a
 * solve run back to back (what the per-kernel timers of ddpb_phase_times assume). */
int ddpb_set_subbatches(ddpb_batch* b, int n);

#ifdef __cplusplus
}
#endif
#endif /* DDPB_H_ */
