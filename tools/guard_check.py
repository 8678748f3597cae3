"""Does the guard limit change which problems are frozen? (development aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
for name in sys.argv[1:]:
    res = {}
    for guard in (3, 10, 30):
        w = synthetic.workload(name, B=1024 if name != "c2" else 4096)
        w.max_restarts = guard
        g = w.generate()
        bt = capi.Batch.from_workload(w)
        bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
        sc = bt.solve(w.iterations, w.mu)
        xs, _ = bt.get_trajectory()
        res[guard] = (sc["status"].copy(), xs.copy(), sc["restarts"].copy())
        print(f"{w.name} guard {guard}: status {np.bincount(sc['status'].astype(int), minlength=3)}, restarts/problem (active) {sc['restarts'][sc['status'] == 0].mean():.2f}, max {sc['restarts'].max():.0f}", flush=True)
        bt.close()
    s3, x3, _ = res[3]
    for guard in (10, 30):
        sg, xg, _ = res[guard]
        both = (s3 == 0) & (sg == 0)
        print(f"   guard 3 vs {guard}: frozen sets equal: {np.array_equal(s3 == 2, sg == 2)}; frozen only under 3: {int(((s3 == 2) & (sg != 2)).sum())}; "
              f"trajectories of the problems active under both identical: {np.array_equal(x3[both], xg[both])}")
