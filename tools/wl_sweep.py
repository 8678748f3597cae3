"""Workload well-posedness sweep (development aid): fraction of problems the guard freezes, restarts and rollouts for
variations of a workload's parameters."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
for off, mu, guard in itertools.product([0.25, 0.1, 0.03], [1e2, 1e4, 1e6], [10, 3]):
    w = synthetic.workload(name, B=B)
    w.target_offset = off; w.mu = mu; w.max_restarts = guard
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    sc = bt.solve(w.iterations, w.mu)
    st = np.bincount(sc["status"].astype(int), minlength=3)
    act = sc["status"] == 0
    print(f"{w.name} offset {off} mu {mu:g} guard {guard}: status {st}, iterations mean {sc['iterations'].mean():.2f}, restarts/problem {sc['restarts'][act].mean():.2f}, "
          f"rollouts/iter {sc['rollouts'][act].mean() / w.iterations:.2f}, opt_constr median {np.median(sc['opt_constr'][act]):.3g} opt_obj median {np.median(sc['opt_obj'][act]):.3g}", flush=True)
    bt.close()
