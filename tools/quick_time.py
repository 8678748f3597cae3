"""Quick per-phase timing of a fixed-iteration solve on one GPU (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic

for name, B in [(a.split(":")[0], int(a.split(":")[1]) if ":" in a else None) for a in (sys.argv[1:] or ["c2", "c3"])]:
    w = synthetic.workload(name, B=B)
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"])
    for rep in range(2):
        bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
        bt.enable_phase_timing(True)
        t0 = time.time(); sc = bt.solve(w.iterations, w.mu); wall = time.time() - t0
        ph = bt.phase_times()
    its = w.iterations
    print(f"{w.name}: B={w.B} T={w.T} iters={its} wall {wall*1e3:.1f} ms  phases(ms) {ph}  rollouts/iter {sc['rollouts'].mean()/its:.2f} "
          f"restarts {sc['restarts'].sum():.0f} knots/s {w.B*w.T*its/wall:.3e} launches {bt.launch_count()}", flush=True)
    bt.close()
