"""Which problems of a workload send the backward pass into long restart loops / line-search give-ups?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
np.set_printoptions(precision=3, linewidth=200)
for name in sys.argv[1:]:
    w = synthetic.workload(name.split(":")[0], B=int(name.split(":")[1]) if ":" in name else None)
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    sc = bt.solve(w.iterations, w.mu)
    rs, ro = sc["restarts"], sc["rollouts"]
    print(w.name, "restarts: mean %.2f" % rs.mean(), "hist", np.histogram(rs, bins=[0, 1, 2, 3, 5, 10, 50, 200, 1e9])[0],
          "| rollouts hist", np.histogram(ro, bins=[0, 11, 15, 30, 100, 1e9])[0], "| nan traj", int(np.isnan(sc["merit"]).sum()),
          "| mu inf", int(np.isinf(sc["mu"]).sum()))
    worst = np.argsort(-rs)[:6]
    for b in worst:
        print("   b=%d restarts %d rollouts %d mu %.3g q0 %s target-q0 %s" % (b, rs[b], ro[b], sc["mu"][b], g["x0"][b, :w.n], (g["targets"][b, :w.n] - g["x0"][b, :w.n]) if w.eq_kind == 1 else ""))
    bt.close()
