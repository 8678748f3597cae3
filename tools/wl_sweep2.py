import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
for dt, off, mu in itertools.product([0.01, 0.005, 0.0025], [0.25, 1.0], [1e4]):
    w = synthetic.workload(name, B=B)
    w.target_offset = off; w.mu = mu; w.dt = dt
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    sc = bt.solve(w.iterations, w.mu)
    st = np.bincount(sc["status"].astype(int), minlength=3)
    act = sc["status"] == 0
    print(f"{w.name} dt {dt} offset {off} mu {mu:g}: status {st}, iterations mean {sc['iterations'].mean():.2f}, restarts/problem {sc['restarts'][act].mean():.2f}, "
          f"rollouts/iter {sc['rollouts'][act].mean() / w.iterations:.2f}, opt_constr median {np.median(sc['opt_constr'][act]):.3g} opt_obj median {np.median(sc['opt_obj'][act]):.3g}", flush=True)
    bt.close()
