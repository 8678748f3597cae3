"""Per-phase cycle breakdown of the linearization kernel (development aid; uses the undeclared debug entry point).
Needs a library built with the counters: DDPB_NVCC_EXTRA=-DDDPB_PHASE_CYCLES=1 python -c "import __graft_entry__ as g; g.build_cuda(force=True)"."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
for name in sys.argv[1:] or ["c3"]:
    w = synthetic.workload(name)
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    bt.compute_derivatives()
    out = (C.c_ulonglong * 16)()
    capi.lib().ddpb_debug_phase_cycles(bt._h, 1, None)
    bt.compute_derivatives()
    capi.lib().ddpb_debug_phase_cycles(bt._h, 1, out)
    v = np.array(list(out), dtype=float)
    names = ["stage", "q_task(last sub-phase)", "v_task(last sub-phase)", "columns", "products", "eq_values", "eq_jac", "assemble",
             "q:local_rot", "q:chain", "q:body", "v:prefix", "v:bias", "v:sweeps", "v:force"]
    tot = v[:15].sum()
    print(w.name, " ".join(f"{n}={100*x/tot:.1f}%" for n, x in zip(names, v[:15])), f"| cycles/CTA {tot / (w.B * w.T / max(1, int(os.environ.get('DDPB_LIN_KPB', '1')))):.0f}")
    bt.close()
