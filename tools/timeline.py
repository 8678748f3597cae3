"""Timeline of one solve (development aid): every launch group with its start / end time on the device (phase timing mode:
direct enqueue on the sub-batch streams), summarised as busy time per phase, the critical chain of one sub-batch and the
overlap achieved."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ddpb_loader
ddpb_loader.load()
from ddp_pinocchio_b200 import capi, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = synthetic.workload(name, B=1024 if name == "c5" else None)
g = w.generate()
bt = capi.Batch.from_workload(w)
bt.set_subbatches(nsub)
bt.set_targets(g["targets"])
for rep in range(2):
    bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    bt.enable_phase_timing(rep == 1)
    bt.solve(w.iterations, w.mu)
buf = (C.c_float * (3 * 20000))()
n = capi.lib().ddpb_debug_timeline(bt._h, buf, 20000)
a = np.array(list(buf[: 3 * n])).reshape(n, 3)
names = ["linearize", "backward", "forward", "other"]
total = a[:, 2].max()
print(f"{w.name} nsub={nsub}: {n} launch groups, solve {total:.2f} ms")
for p in range(4):
    m = a[:, 0] == p
    d = a[m, 2] - a[m, 1]
    print(f"  {names[p]:10s} n={m.sum():4d} sum {d.sum():8.2f} ms  mean {d.mean():.3f}  max {d.max():.3f}  min {d.min():.3f}")
# concurrency profile: how many launch groups are in flight over time
ts = np.linspace(0, total, 2000)
inflight = ((a[:, 1][None, :] <= ts[:, None]) & (a[:, 2][None, :] > ts[:, None])).sum(1)
print("  in-flight launch groups: mean %.2f, fraction of time with 0: %.3f, with 1: %.3f" % (inflight.mean(), (inflight == 0).mean(), (inflight == 1).mean()))
if len(sys.argv) > 3:
    for row in a[: int(sys.argv[3])]:
        print("   %-10s %8.3f -> %8.3f  (%.3f)" % (names[int(row[0])], row[1], row[2], row[2] - row[1]))
bt.close()
