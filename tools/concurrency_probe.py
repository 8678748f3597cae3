"""Do two independent batches on their own streams overlap? (development aid)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
def mk(first):
    w = synthetic.workload(name, B=B)
    g = w.generate(first_problem=first)
    bt = capi.Batch.from_workload(w)
    bt.set_subbatches(1)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    bt.snapshot_save()
    return w, bt
w, a = mk(0); _, b = mk(B)
def run(bts, reps=3):
    best = 1e9
    for _ in range(reps):
        for bt in bts: bt.snapshot_restore()
        for bt in bts: bt.sync()
        t0 = time.perf_counter()
        for bt in bts: bt.solve(w.iterations, w.mu, read_back=False)
        for bt in bts: bt.sync()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3
print(f"{name} B={B}: one batch alone {run([a]):.1f} ms; two batches concurrently {run([a, b]):.1f} ms (sum if serialized: {2*run([a]):.1f})")
w2 = synthetic.workload(name, B=2 * B); g2 = w2.generate(); c = capi.Batch.from_workload(w2); c.set_subbatches(1)
c.set_targets(g2["targets"]); c.make_trajectory(g2["x0"], g2["us"]); c.set_multipliers(None, g2["mult_jac"]); c.snapshot_save()
print(f"   one batch of {2*B}: {run([c]):.1f} ms")
