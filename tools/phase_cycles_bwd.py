"""Per-phase cycle breakdown of the backward kernel (development aid; uses the undeclared debug entry point).
Needs a library built with the counters: DDPB_NVCC_EXTRA=-DDDPB_PHASE_CYCLES=1 python -c "import __graft_entry__ as g; g.build_cuda(force=True)"."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
for name in sys.argv[1:] or ["c3"]:
    w = synthetic.workload(name)
    if name == "c5":
        w.B = 1024
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    bt.compute_derivatives()
    out = (C.c_ulonglong * 16)()
    mu, reg = bt.backward_pass(1e7, 1.0)   # large mu / reg: no restarts, one clean pass
    capi.lib().ddpb_debug_phase_cycles(bt._h, 1, None)
    bt.set_scalar("status", 0.0)
    mu, reg = bt.backward_pass(1e7, 1.0)
    capi.lib().ddpb_debug_phase_cycles(bt._h, 1, out)
    v = np.array(list(out), dtype=float)[8:16]
    names = ["A4 prefetch-issue+sync", "B VF/VU", "C Q", "D chol+solve", "F V update", "A1 setup+head", "A2 cp.async wait", "A3 contract"]
    tot = v.sum()
    print(w.name, " ".join(f"{n}={100*x/tot:.1f}%" for n, x in zip(names, v)), f"| cycles/knot {tot / (w.B * w.T):.0f}", "restarts", bt.get_scalars()["restarts"].sum())
    bt.close()
