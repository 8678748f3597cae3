// model_host.hpp -- host-side validation / conversion of a ddpb_model_desc into the POD the kernels
// take by value (replaces pinocchio::urdf::buildModel + ModelTpl, pinocchio_model.ipp:98-152, and
// pendulum_model_t's constructor asserts, pendulum_model.hpp:29-33).  Pure C++ (no CUDA calls).
#pragma once
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/ddpb.h"
#include "rbd.cuh"

namespace ddpb {

// returns DDPB_OK or an error code with a message in `err`
inline int model_from_desc(const ddpb_model_desc* d, ModelDev& M, std::string& err) {
  std::memset(&M, 0, sizeof(ModelDev));
  M.kind = d->kind;
  M.pend_m = d->pend_mass; M.pend_l = d->pend_length;
  for (int k = 0; k < 3; ++k) M.grav[k] = d->gravity[k];
  if (d->kind == DDPB_MODEL_PENDULUM) {
    if (!(d->pend_mass > 0) || !(d->pend_length > 0)) { err = "mass and length should be positive"; return DDPB_ERR_INVALID; }
    M.n = 1;
    return DDPB_OK;
  }
  if (d->kind != DDPB_MODEL_REVOLUTE_CHAIN) { err = "unknown model kind"; return DDPB_ERR_UNSUPPORTED; }
  const int n = d->n_joints;
  if (n < 1 || n > MAXJ) { err = "n_joints must be 1..7"; return DDPB_ERR_UNSUPPORTED; }
  if (d->n_frames < 0 || d->n_frames > MAXF) { err = "at most 16 frames"; return DDPB_ERR_UNSUPPORTED; }
  if (!d->parent || !d->placement_rot || !d->placement_trans || !d->axis || !d->mass || !d->com || !d->inertia) { err = "null model array"; return DDPB_ERR_INVALID; }
  for (int i = 0; i < n; ++i)
    if (d->parent[i] != i - 1) { err = "only serial chains (parent[i] == i-1) are supported"; return DDPB_ERR_UNSUPPORTED; }
  M.n = n;
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 9; ++k) M.Rp[i][k] = d->placement_rot[i * 9 + k];
    double an = 0;
    for (int k = 0; k < 3; ++k) { M.pp[i][k] = d->placement_trans[i * 3 + k]; M.axis[i][k] = d->axis[i * 3 + k]; an += M.axis[i][k] * M.axis[i][k]; }
    if (std::fabs(an - 1.0) > 1e-9) { err = "joint axes must be unit vectors"; return DDPB_ERR_INVALID; }
    const int jt = d->joint_type ? d->joint_type[i] : DDPB_JOINT_REVOLUTE;
    if (jt != DDPB_JOINT_REVOLUTE && jt != DDPB_JOINT_PRISMATIC) { err = "joint types must be revolute (0) or prismatic (1)"; return DDPB_ERR_UNSUPPORTED; }
    M.jtype[i] = jt;
    if (jt != DDPB_JOINT_REVOLUTE) M.any_prismatic = 1;
    if (jt != DDPB_JOINT_REVOLUTE) M.any_prismatic = 1;
    const double ms = d->mass[i];
    const double* c = d->com + 3 * i;
    const double* I = d->inertia + 6 * i;
    if (!(ms > 0)) { err = "link masses must be positive"; return DDPB_ERR_INVALID; }
    // inertia about the body-frame origin: Io = Ic + m (|c|^2 1 - c c^T)
    const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    M.Io[i][0] = I[0] + ms * (cc - c[0] * c[0]);
    M.Io[i][1] = I[1] - ms * c[0] * c[1];
    M.Io[i][2] = I[2] - ms * c[0] * c[2];
    M.Io[i][3] = I[3] + ms * (cc - c[1] * c[1]);
    M.Io[i][4] = I[4] - ms * c[1] * c[2];
    M.Io[i][5] = I[5] + ms * (cc - c[2] * c[2]);
    for (int k = 0; k < 3; ++k) M.h[i][k] = ms * c[k];
    M.mass[i] = ms;
  }
  if (d->n_frames > 0 && (!d->frame_parent || !d->frame_trans)) { err = "null frame array"; return DDPB_ERR_INVALID; }
  M.nframes = d->n_frames;
  for (int f = 0; f < d->n_frames; ++f) {
    if (d->frame_parent[f] < 0 || d->frame_parent[f] >= n) { err = "frame parent out of range"; return DDPB_ERR_INVALID; }
    M.fparent[f] = d->frame_parent[f];
    for (int k = 0; k < 3; ++k) M.fp[f][k] = d->frame_trans[f * 3 + k];
  }
  return DDPB_OK;
}

}  // namespace ddpb
