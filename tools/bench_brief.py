"""One-line summary of a bench.py JSON line read from stdin (development aid)."""
import sys, json
for ln in sys.stdin.read().strip().splitlines():
    try:
        d = json.loads(ln)
    except Exception:
        continue
    k = {a: round(v["ms_avg"], 3) for a, v in d.get("kernels", {}).items()}
    print(d["config"]["workload"], "n_gpus", d["n_gpus"], "value %.3e ms/step %.2f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]),
          "one-stream %.2f" % d.get("kernel_timing", {}).get("ms_per_step_one_stream", float("nan")), k,
          "roofline", {a: (round(b, 3) if isinstance(b, float) else b) for a, b in d.get("roofline", {}).items() if a in ("kernel", "frac", "achieved")},
          "status", d.get("status"))
