"""Backward-pass timing (development aid): a pass as it occurs in a solve (with its restarts) and the same pass repeated
with the mu / reg it ended with (clean: no restarts), plus the distribution of restarts over problems."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
for arg in sys.argv[1:] or ["c2", "c3"]:
    name, B = (arg.split(":") + [None])[:2]
    w = synthetic.workload(name, B=int(B) if B else None)
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    for iters in (0, 3, 8):
        bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
        sc = bt.solve(iters, w.mu) if iters else None
        mu = sc["mu"] if sc else np.full(w.B, w.mu)
        reg = sc["reg"] if sc else np.zeros(w.B)
        bt.set_scalar("status", 0.0)
        bt.compute_derivatives()
        r0 = bt.get_scalars()["restarts"].copy()
        bt.enable_phase_timing(True)
        mu1, reg1 = bt.backward_pass(mu, reg)
        t_real = bt.phase_times()["backward"]
        rr = bt.get_scalars()["restarts"] - r0
        st = bt.get_scalars()["status"]
        r1 = bt.get_scalars()["restarts"].copy()
        bt.backward_pass(mu1, reg1)   # the guarded problems stay frozen: a pass without a single restart
        t_clean = bt.phase_times()["backward"]
        assert (bt.get_scalars()["restarts"] == r1).all()
        bt.compute_derivatives(); t_lin = bt.phase_times()["linearize"]
        ro0 = bt.get_scalars()["rollouts"].copy()
        step = bt.forward_pass(); t_fwd = bt.phase_times()["forward"]
        act = bt.get_scalars()["status"] == 0
        trials = (bt.get_scalars()["rollouts"] - ro0)[act]
        print("   line search: trials needed per problem, histogram 1,2,3,4,5-8,9-16,17-34:", np.histogram(trials, bins=[1, 2, 3, 4, 5, 9, 17, 35])[0], "max", trials.max())
        bt.enable_phase_timing(False)
        print(f"{w.name} B={w.B} after {iters} iterations: backward {t_real:.3f} ms with restarts (total {rr.sum():.0f}, max/problem {rr.max():.0f}, "
              f"problems {int((rr > 0).sum())}, guarded {int((st == 2).sum())}), clean {t_clean:.3f} ms = {t_clean * 1e3 / w.T:.2f} us/knot; "
              f"linearize {t_lin:.3f} ms; forward {t_fwd:.3f} ms", flush=True)
    bt.close()
