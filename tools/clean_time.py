"""Clean-pass timings (development aid): backward pass with mu / reg large enough that nothing restarts, forward pass
right after it (full step accepted almost everywhere), linearization; whole batch in one launch each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
for arg in sys.argv[1:] or ["c2", "c3"]:
    name, B = (arg.split(":") + [None])[:2]
    w = synthetic.workload(name, B=int(B) if B else (1024 if name == "c5" else None))
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    bt.compute_derivatives()
    bt.backward_pass(1e7, 1.0)
    for rep in range(2):
        bt.set_scalar("status", 0.0); bt.set_scalar("restarts", 0.0); bt.set_scalar("rollouts", 0.0)
        bt.enable_phase_timing(True)
        bt.backward_pass(1e7, 1.0)
        tb = bt.phase_times()["backward"]
        rs = bt.get_scalars()["restarts"].sum()
        bt.set_scalar("mu", 1e7)
        bt.forward_pass()
        tf = bt.phase_times()["forward"]
        ro = bt.get_scalars()["rollouts"]
        bt.compute_derivatives()
        tl = bt.phase_times()["linearize"]
        bt.enable_phase_timing(False)
    print(f"{w.name} B={w.B}: clean backward {tb:.3f} ms = {tb * 1e3 / w.T:.2f} us/knot (restarts {rs:.0f}); forward {tf:.3f} ms = {tf * 1e3 / w.T:.2f} us/knot "
          f"(rollouts max {ro.max():.0f} mean {ro.mean():.2f}); linearize {tl:.3f} ms", flush=True)
    bt.close()
