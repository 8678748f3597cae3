"""Small solves of every workload family for compute-sanitizer (development aid):
   compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_run.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
for name, B, T in [("c1", 3, 12), ("c2", 70, 16), ("c3", 36, 10), ("c4", 33, 10), ("c5", 9, 8)]:
    for nsub in (1, 2):
        w = synthetic.workload(name, B=B, T=T)
        g = w.generate()
        bt = capi.Batch.from_workload(w)
        bt.set_subbatches(nsub)
        bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
        sc = bt.solve(3, w.mu)
        xs, us = bt.get_trajectory()
        print(name, nsub, "status", np.bincount(sc["status"].astype(int), minlength=3), "finite", bool(np.isfinite(xs[sc["status"] == 0]).all()), flush=True)
        bt.close()
