"""Backward-pass time versus batch size (development aid): separates per-problem latency from throughput limits."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
for B in [int(a) for a in sys.argv[2:]] or [148, 296, 592, 1024, 2048]:
    w = synthetic.workload(name, B=B)
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    sc = bt.solve(3, w.mu)
    bt.compute_derivatives()
    r0 = bt.get_scalars()["restarts"].copy()
    bt.enable_phase_timing(True)
    bt.backward_pass(sc["mu"], sc["reg"])
    ph = bt.phase_times()
    rr = bt.get_scalars()["restarts"] - r0
    r1, r0 = rr.sum(), 0
    print("   restarts in the timed pass: total %d, max per problem %d, problems with any %d" % (rr.sum(), rr.max(), (rr > 0).sum()))
    bt.compute_derivatives()
    ph2 = bt.phase_times()
    bt.forward_pass()
    ph3 = bt.phase_times()
    print(f"{w.name} B={B}: backward {ph['backward']:.3f} ms ({ph['backward']*1e3/w.T:.2f} us/knot, restarts in pass {r1-r0:.0f}); linearize {ph2['linearize']:.3f} ms; forward {ph3['forward']:.3f} ms", flush=True)
    bt.close()
