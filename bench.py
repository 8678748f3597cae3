#!/usr/bin/env python
"""bench.py -- throughput of the batched DDP hot path (linearize + backward pass + line-search rollouts).

One "step" = one complete fixed-iteration DDP solve (ddp_solver_t::solve, ddp.hpp:744-842, `iterations`
iterations + the reference's warm-up pass) of the whole synthetic batch from its initial trajectory.
  value : knots linearized/s = (sum over problems of the iterations carried through) * T * steps / time -- a problem the
          guard froze (status 2) or that converged early stops counting -- inputs already resident in HBM (a device-side
          snapshot is restored at the start of every step), CUDA events on the launching stream;
  e2e   : the same metric through the C ABI with HOST (pinned) inputs: x0 / controls / targets / multiplier
          seeds go host->device, the initial rollout runs, the solve runs, and per-problem scalars plus the
          final trajectory come back device->host, every step.
N > 1: one process per GPU (torchrun), independent problem shards (weak scaling), NCCL only gathers the
per-problem scalars.  `--impl reference` times the CPU oracle restatement of the reference path instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ddpb_loader  # noqa: E402

ddpb_loader.load()
from ddp_pinocchio_b200 import synthetic  # noqa: E402

METRIC = "DDP iterations/sec (batch x horizon knots linearized/sec)"


def algorithmic_bytes_per_knot(w):
    """BASELINE.md section 3 / SURVEY.md 8(d): bytes one linearization must write / one backward step must read, per knot,
    counted on the reference's dense derivative_storage_t."""
    nd, nu, nx = w.ndx, w.nu, w.nx
    ne = w.ne_per_knot().astype(np.int64)
    slab = (nd + nu + nd * nd + nu * nd + nu * nu) + (nx + nd * nd + nd * nu + nd ** 3 + nd * nd * nu + nd * nu * nu) + \
        ne * (1 + nd + nu + nd * nd + nu * nd + nu * nu)
    lin = 8.0 * (slab + nx + nu)
    bwd = 8.0 * (slab + 2 * nx + ne + ne * nd + nu + nu * nd)
    return float(lin.mean()), float(bwd.mean())


def stored_bytes_per_knot(w):
    """The same two figures for the records the kernels actually move (csrc/batch.cuh: the constants of the reference's
    cost / Euler scheme and, with FD-of-Jacobians Hessians, the structurally zero tensor entries are not stored)."""
    n, nd, nu, nx = w.n, w.ndx, w.nu, w.nx
    ne = w.ne_per_knot().astype(np.int64)
    dense = w.hess_mode == synthetic.HESS_FD_VALUES
    tens = (nd ** 3 + nd * nu * nd + nd * nu * nu) if dense else (n * nd * nd + n * nu * n)
    rec = nx + n * nd + n * nu + nu + tens + ne * (1 + nd + nu + nd * nd + nu * nd + nu * nu)
    rec = (rec + 1) // 2 * 2
    lin = 8.0 * (rec + nx + nu)
    bwd = 8.0 * (rec - nx + 2 * nx + ne + ne * nd + nu + nu * nd)   # the backward pass does not read f
    return float(lin.mean()), float(bwd.mean())


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons; samples are time-stamped on arrival so that only those that fall inside
    the timed region are reported."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def count(self, t0, t1=None):
        return sum(1 for t, _ in self.rows if t >= t0 and (t1 is None or t <= t1))

    def stop(self, windows):
        """windows: list of (t0, t1) wall-clock intervals during which the timed kernels were running"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def kernel_counts(workload_name):
    """Per-kernel figures taken from ncu captures of this workload (profiles/r02_kernel_counts.json): executed FP64
    flops per knot and DRAM bytes per knot (dram__bytes_read.sum + dram__bytes_write.sum of one launch / its knots)."""
    for name in ("r02_kernel_counts.json",):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                return json.load(f).get(workload_name, {})
    return {}


def calibrate_cpu_sample(ol, w, nthreads, target_s):
    """Number of problems the CPU path solves in about `target_s` seconds of wall time on `nthreads` threads."""
    g = w.generate(B=nthreads)
    sec, _ = ol.solve_batch(w, g, nthreads, nthreads, w.iterations, w.mu)
    n = int(max(nthreads, min(w.B, nthreads * round(target_s / max(sec, 1e-3)))))
    return n


def run_reference(args, w, rank):
    """The reference's own CPU path (restated by oracle/; the reference cannot be compiled here)
    on all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    import oracle_lib as ol
    nthreads = os.cpu_count() or 1
    nprob = args.cpu_sample or calibrate_cpu_sample(ol, w, nthreads, 5.0)
    g = w.generate(B=nprob)
    for _ in range(args.warmup):
        ol.solve_batch(w, g, min(nprob, nthreads), nthreads, 1, w.mu)
    t = 0.0
    ro = rs = its_done = div = 0
    for _ in range(args.steps):
        sec, info = ol.solve_batch(w, g, nprob, nthreads, w.iterations, w.mu)
        t += sec; ro += info["rollouts"]; rs += info["restarts"]; its_done += info["iterations"]; div = info["diverged"]
    val = its_done * w.T / t   # knots of the iterations carried through (a problem frozen by the guard stops counting)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "knots/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "ddp_iterations_per_sec": its_done / t,
            "config": {"workload": w.name, "problems": nprob, "horizon": w.T, "iterations_per_step": w.iterations,
                       "counting": "problem-iterations carried through x horizon (frozen problems stop counting)"},
            "status": {"problems": nprob, "guarded": int(div)},
            "cpu_baseline": {"value": val, "unit": "knots/s", "cores": nthreads, "kind": "port",
                             "sample": f"{nprob} of {w.B} problems x {w.T} knots x {w.iterations} iterations per step, OpenMP over problems"},
            "e2e": {"value": val, "unit": "knots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rollouts_per_iteration": ro / (nprob * w.iterations * args.steps), "backward_restarts_per_problem": rs / (nprob * args.steps)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", help="c1..c5 (BASELINE.json configs); default configs[1]")
    ap.add_argument("--batch", type=int, default=None, help="problems per GPU (default: the config's batch)")
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    w = synthetic.workload(args.workload)
    if args.batch:
        w.B = args.batch
    elif args.workload == "c5":
        w.B = 1024  # 8192 problems over 8 GPUs = 1024 per GPU
    if args.impl == "reference":
        run_reference(args, w, rank)
        return

    import torch
    import torch.distributed as dist
    from ddp_pinocchio_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, T, its = w.B, w.T, w.iterations
    g = w.generate(first_problem=rank * B)  # weak scaling: every rank owns its own B problems
    bt = capi.Batch.from_workload(w, device=local)
    stream = torch.cuda.Stream(device=local)
    bt.set_stream(stream.cuda_stream)
    # pinned host inputs / outputs for the e2e leg
    pins = []

    def pin(a):
        arr, p = capi.host_alloc(a.nbytes)
        pins.append(p)
        arr = arr.reshape(a.shape)
        arr[...] = a
        return arr
    h_x0, h_us, h_tg, h_mj = pin(g["x0"]), pin(g["us"]), pin(g["targets"]) if g["targets"].size else g["targets"], \
        pin(g["mult_jac"]) if g["mult_jac"].size else g["mult_jac"]
    h2d = h_x0.nbytes + h_us.nbytes + h_tg.nbytes + h_mj.nbytes
    d2h = (capi.NSCALARS * B + B * (T + 1) * w.nx + B * T * w.nu) * 8

    # results land in pinned host buffers too (a D2H copy into pageable memory is staged by the driver at a fraction of the link rate)
    h_xs = pin(np.zeros((B, T + 1, w.nx))); h_uso = pin(np.zeros((B, T, w.nu))); h_sc = pin(np.zeros((capi.NSCALARS, B)))

    def e2e_step():
        bt.set_targets(h_tg)
        bt.make_trajectory(h_x0, h_us)
        bt.set_multipliers(None, h_mj)
        sc = bt.solve(its, w.mu, out=h_sc)
        xs, us = bt.get_trajectory(out=(h_xs, h_uso))
        return sc, xs

    # resident leg: snapshot once, restore per step
    bt.set_targets(h_tg); bt.make_trajectory(h_x0, h_us); bt.set_multipliers(None, h_mj)
    bt.snapshot_save()
    scal_ptr = bt.scalars_device_ptr()

    class _Holder:
        __cuda_array_interface__ = {"shape": (capi.NSCALARS * B,), "typestr": "<f8", "data": (scal_ptr, False), "version": 3}
    scal_t = torch.as_tensor(_Holder(), device=torch.device("cuda", local))
    gathered = torch.empty(world * capi.NSCALARS * B, dtype=torch.float64, device=scal_t.device) if world > 1 else None

    def resident_step():
        bt.snapshot_restore()
        bt.solve(its, w.mu, read_back=False)
        if world > 1:  # the path's only exchange: per-problem cost / merit / convergence scalars
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, scal_t)

    def barrier():
        bt.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        resident_step()
    barrier()
    # ---- timed region: K steps, CUDA events on the launching stream (the solver forks its sub-batch streams from it
    # and joins them back into it), barrier + synchronize on both sides
    l0 = bt.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    windows = []
    tw0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        resident_step()
    ev1.record(stream)
    barrier()
    windows.append((tw0, time.time()))
    launches = bt.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    sc = bt.get_scalars()
    if rank == 0:
        # a short timed region can end before nvidia-smi delivered enough samples: keep the same steps running (untimed)
        # until a handful of samples under this load exist
        t_ext = time.time()
        while sampler.proc and sum(sampler.count(a, b) for a, b in windows) + sampler.count(t_ext) < 5 and time.time() - t_ext < 4.0:
            bt.snapshot_restore(); bt.solve(its, w.mu, read_back=False); bt.sync()   # no collective: rank 0 only
        if time.time() - t_ext > 0.05:
            windows.append((t_ext, time.time()))
    barrier()
    clocks = sampler.stop(windows) if rank == 0 else None
    if rank == 0 and len(windows) > 1:
        clocks["note"] = "timed region shorter than the sampling period: sampling continued over identical untimed steps"
    # ---- per-kernel launch durations: same steps with the solver on ONE stream (kernels back to back, no overlap), one
    # CUDA-event pair around every launch
    ksteps = min(args.steps, 3)
    bt.set_subbatches(1)
    resident_step(); barrier()
    bt.enable_phase_timing(True)
    evk0, evk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evk0.record(stream)
    for _ in range(ksteps):
        resident_step()
    evk1.record(stream)
    barrier()
    ms_one_stream = evk0.elapsed_time(evk1) / ksteps
    phases = bt.phase_times()
    bt.enable_phase_timing(False)
    bt.set_subbatches(int(os.environ.get("DDPB_SUBBATCHES", "0")) or bt.default_subbatches())
    # e2e leg (host buffers in, results out), wall clock bracketed by synchronisation
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sc_e2e, _ = e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        tt = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=scal_t.device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tt[0]), float(tt[1])
    # problem-iterations carried through in one solve (identical every step: same inputs), summed over the ranks
    its_done = float(sc["iterations"].sum())
    st_hist = [int((sc["status"] == k).sum()) for k in (0, 1, 2)]
    if world > 1:
        tt = torch.tensor([its_done] + st_hist, dtype=torch.float64, device=scal_t.device)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        its_done = float(tt[0]); st_hist = [int(v) for v in tt[1:].tolist()]
    if rank == 0:
        knots_per_step = its_done * T
        value = knots_per_step * args.steps / (ms * 1e-3)
        e2e_val = knots_per_step * args.steps / (e2e_ms * 1e-3)
        hbm_peak, peak_src = read_peaks()
        lin_b, bwd_b = algorithmic_bytes_per_knot(w)
        lin_sb, bwd_sb = stored_bytes_per_knot(w)
        n_lin = n_bwd = (its + 1) * ksteps  # launches in the kernel-timing leg (incl. the warm-up pass of every solve)
        kc = kernel_counts(w.name)
        kern = {
            "linearize": {"ms_avg": phases["linearize"] / n_lin, "alg_bytes": lin_b * B * T, "stored_bytes": lin_sb * B * T},
            "backward": {"ms_avg": phases["backward"] / n_bwd, "alg_bytes": bwd_b * B * T, "stored_bytes": bwd_sb * B * T},
            "forward": {"ms_avg": phases["forward"] / n_bwd, "alg_bytes": None, "stored_bytes": None},
        }
        dom = max(("linearize", "backward", "forward"), key=lambda k: phases[k])
        share = {k: phases[k] / max(sum(phases.values()), 1e-9) for k in phases}
        rname = dom if kern[dom]["alg_bytes"] else "linearize"
        rk = kern[rname]
        achieved = rk["alg_bytes"] / (rk["ms_avg"] * 1e-3) / 1e9
        fp64_peak = capi.measure_fp64_peak(local)
        traffic = kc.get(rname, {}).get("dram_bytes_per_knot")
        kout = {}
        for k, v in kern.items():
            kout[k] = {"ms_avg": v["ms_avg"], "share_of_step": share[k],
                       "hbm_gbs": (v["alg_bytes"] / (v["ms_avg"] * 1e-3) / 1e9) if v["alg_bytes"] else None,
                       "hbm_gbs_stored_records": (v["stored_bytes"] / (v["ms_avg"] * 1e-3) / 1e9) if v["stored_bytes"] else None}
            fl = kc.get(k, {}).get("fp64_flops_per_knot")
            if fl:
                kout[k]["fp64_tflops"] = fl * B * T / (v["ms_avg"] * 1e-3) / 1e12
                kout[k]["fp64_frac_of_measured_peak"] = kout[k]["fp64_tflops"] / fp64_peak
        line = {
            "metric": METRIC, "value": value, "unit": "knots/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "ddp_iterations_per_sec": its_done * args.steps / (ms * 1e-3),
            "config": {"workload": w.name, "problems_per_gpu": B, "horizon": T, "iterations_per_step": its, "nx": w.nx, "nu": w.nu,
                       "hess_mode": "fd_of_jacobians" if w.hess_mode == 0 else "fd_of_values", "mu": w.mu,
                       "l2": "working set (derivative slabs %.2f GB per GPU) is larger than L2" % (lin_sb * B * T / 1e9),
                       "counting": "problem-iterations carried through x horizon (frozen problems stop counting)",
                       "parallelism": f"dp{world} (independent problem shards, NCCL gathers per-problem scalars only)"},
            "e2e": {"value": e2e_val, "unit": "knots/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                    "ms_per_step": e2e_ms / args.steps, "timer": "wall clock between synchronisations"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": rname, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": (traffic * B * T) if traffic else None, "peak_source": peak_src,
                         "achieved_stored_records": rk["stored_bytes"] / (rk["ms_avg"] * 1e-3) / 1e9,
                         "note": "algorithmic bytes per launch (BASELINE.md s3 / SURVEY 8d, counted on the reference's dense "
                                 "derivative_storage_t: 8*(slab+nx+nu) per knot written by the linearization, 8*(slab+2nx+nu+nu*ndx+eq) read "
                                 "by the backward pass) / mean CUDA-event launch duration; the kernels move only the state-dependent "
                                 "entries (csrc/batch.cuh) and achieved_stored_records is the rate on the bytes actually stored; traffic "
                                 "(ncu dram bytes per knot, profiles/r02_kernel_counts.json, x knots per launch) is " +
                                 ("BELOW the algorithmic bytes for that reason" if not traffic or traffic * B * T <= rk["alg_bytes"] else
                                  "%.2fx the algorithmic bytes: the value-based mode stores dense tensors and its kernel (255 registers) spills "
                                  "to local memory, whose lines are evicted from L2 by the streamed records (DESIGN.md 4.1)"
                                  % (traffic * B * T / rk["alg_bytes"]))},
            "roofline_fp64": ({"kernel": "linearize", "bound": "fp64", "achieved": kout["linearize"]["fp64_tflops"], "peak": fp64_peak, "unit": "TFLOP/s",
                               "frac": kout["linearize"]["fp64_frac_of_measured_peak"],
                               "note": "executed FP64 flops per knot (ncu, profiles/r02_kernel_counts.json: 2*dfma + dadd + dmul) x knots per "
                                       "pass / mean duration of a linearization pass; peak = ddpb_measure_fp64_peak on this GPU"}
                              if "fp64_tflops" in kout["linearize"] else None),
            "kernels": kout,
            "kernel_timing": {"steps": ksteps, "ms_per_step_one_stream": ms_one_stream,
                              "note": "launch durations measured with the solver on one stream (no overlap between sub-batches)"},
            "fp64_peak_tflops_measured": fp64_peak,
            "status": {"problems": world * B, "active": st_hist[0], "converged": st_hist[1], "guarded": st_hist[2],
                       "problem_iterations_per_solve": its_done},
            "rollouts_per_iteration": float(sc["rollouts"].mean() / its), "backward_restarts_per_problem": float(sc["restarts"].mean()),
            "opt_constr_median": float(np.median(sc["opt_constr"])),
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle_lib as ol
            nthreads = os.cpu_count() or 1
            nprob = args.cpu_sample or calibrate_cpu_sample(ol, w, nthreads, 10.0)
            gs = w.generate(B=nprob)
            sec, info = ol.solve_batch(w, gs, nprob, nthreads, its, w.mu)
            line["cpu_baseline"] = {"value": info["iterations"] * T / sec, "unit": "knots/s", "cores": nthreads, "kind": "port",
                                    "sample": f"{nprob} of {B} problems x {T} knots x {its} iterations, OpenMP over problems, {sec:.1f} s"}
        print(json.dumps(line), flush=True)
    bt.close()
    for p in pins:
        capi.host_free(p)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
