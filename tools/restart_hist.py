"""Where in the horizon do backward passes fail (development aid; c2-type models: thread-per-problem kernel counters)?"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
for arg in sys.argv[1:] or ["c2"]:
    w = synthetic.workload(arg)
    g = w.generate()
    bt = capi.Batch.from_workload(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    out = (C.c_ulonglong * 16)()
    capi.lib().ddpb_debug_phase_cycles(bt._h, 1, None)
    sc = bt.solve(0, w.mu)   # warm-up pass only
    capi.lib().ddpb_debug_phase_cycles(bt._h, 1, out)
    print(w.name, "warm-up pass failures by horizon sixteenth (0 = knots near t=0):", list(out), flush=True)
    bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    sc = bt.solve(w.iterations, w.mu)
    capi.lib().ddpb_debug_phase_cycles(bt._h, 1, out)
    print(w.name, "whole solve:", list(out), "status hist", np.bincount(sc["status"].astype(int), minlength=3), "restarts (after warm-up)", sc["restarts"].sum(), flush=True)
    bt.enable_phase_timing(True)
    bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    sc = bt.solve(w.iterations, w.mu)
    print("   phase ms", bt.phase_times())
    bt.close()
