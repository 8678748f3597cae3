"""Turn the raw captures of tools/capture_r02.sh (gpurun_out/r02_*) into the committed summaries under profiles/.

  profiles/r02_ncu_full_<workload>.txt   selected `ncu --set full` metrics, stall-reason split and the hottest source lines of
                                          the main launch of every hot kernel (solver on one stream)
  profiles/r02_kernel_metrics.txt        flops / DRAM bytes / duration of the kernels of tools/bwd_scan.py (3 iterations + probes)
  profiles/r02_kernel_counts.json        per-knot executed FP64 flops and DRAM bytes (read by bench.py)
  profiles/r02_launches_bench_c2_*.csv   ncu launch lists of the bench command + r02_launch_shares.txt
(the round-1 files profiles/r01_* were produced by the same script at the end of round 1)
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "lts__t_sector_hit_rate.pct"]


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def full_summary(rep, out):
    rows = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    if len(rows) < 3:
        return
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    out.write(f"## {name}\n")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.write(f"  {w:72s} {vals[i]} {units[i]}\n")
    src = list(csv.reader(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]).splitlines()))
    h = None; fname = ""; cur = None; agg = collections.OrderedDict(); stalls = collections.Counter()
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            h = r; continue
        if h is None or len(r) < 10:
            continue
        if r[0] != "":
            cur = (fname, r[0], r[1].strip()[:90]); agg.setdefault(cur, collections.Counter())
        d = dict(zip(h[2:], r[2:]))
        if d.get("Address") and cur is not None:
            agg[cur]["samples"] += num(d["# Samples"]); agg[cur]["inst"] += num(d["Instructions Executed"])
            for k in h:
                if k.startswith("stall_") and "Not Issued" not in k:
                    stalls[k] += num(d.get(k))
    tot = max(sum(a["samples"] for a in agg.values()), 1); st = max(sum(stalls.values()), 1)
    out.write("  warp stall samples: " + ", ".join(f"{k[6:]} {100*v/st:.1f}%" for k, v in stalls.most_common(7)) + "\n")
    out.write("  hottest source lines (share of samples):\n")
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:10]:
        out.write(f"    {100*a['samples']/tot:5.1f}%  {k[0]}:{k[1]}  {k[2]}\n")
    out.write("\n")


def tob(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def main():
    os.makedirs(P, exist_ok=True)
    for wl, label in (("c2", "c2_chain2r_4096x200"), ("c3", "c3_arm6r_1024x100"), ("c5", "c5_arm7r_fd_8192x200, 1024 problems per GPU")):
        reps = [os.path.join(G, f"r02_full_{wl}_{k}.ncu-rep") for k in ("linearize_pt_kernel", "linearize_kernel", "backward_lg_kernel", "backward_kernel", "ls_trials_kernel")]
        reps = [r for r in reps if os.path.exists(r)]
        if not reps:
            continue
        with open(os.path.join(P, f"r02_ncu_full_{wl}.txt"), "w") as out:
            out.write(f"# ncu --set full --clock-control none --import-source on, DDPB_SUBBATCHES=1 python tools/bwd_scan.py {wl}  ({label})\n"
                      "# round 2: main launch of each hot kernel (whole batch in one launch)\n\n")
            for r in reps:
                full_summary(r, out)
    counts = {"_source": "ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum,dram__bytes_{read,write}.sum "
                         "--clock-control none, DDPB_SUBBATCHES=1 python tools/bwd_scan.py <workload> (tools/capture_r02.sh): every hot-kernel "
                         "launch of the script; flops = 2*dfma + dadd + dmul; linearize: sum over all launches / (passes x problems x knots), "
                         "a pass being one launch per run of knots with equal equality rows; backward: the shortest launch (a pass "
                         "without restarts) / (problems x knots)"}
    names = {"c2": ("c2_chain2r_4096x200", 4096, 200), "c3": ("c3_arm6r_1024x100", 1024, 100), "c5": ("c5_arm7r_fd_8192x200", 1024, 200)}
    with open(os.path.join(P, "r02_kernel_metrics.txt"), "w") as out:
        for wl, (label, B, T) in names.items():
            f = os.path.join(G, f"r02_metrics_{wl}.csv")
            if not os.path.exists(f):
                continue
            rows = list(csv.reader(open(f)))
            hdr = None; per = collections.OrderedDict()
            for r in rows:
                if "Kernel Name" in r:
                    hdr = r; continue
                if hdr and len(r) == len(hdr):
                    d = dict(zip(hdr, r))
                    per.setdefault((d["ID"], d["Kernel Name"].split("(")[0].replace("void ", "").replace("ddpb::", ""), d.get("Grid Size", "")), {})[d["Metric Name"]] = \
                        (float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
            out.write(f"== {label} (B={B}, T={T}, one stream)\n")
            counts[label] = {}
            acc = {"linearize": [], "backward": []}
            for k, m in per.items():
                fl = 2 * m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"][0] + m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"][0] + \
                    m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"][0]
                t = m["gpu__time_duration.sum"]; tms = t[0] / 1e6 if t[1] == "ns" else (t[0] / 1e3 if t[1] in ("us", "usecond") else t[0])
                rd = tob(*m["dram__bytes_read.sum"]); wr = tob(*m["dram__bytes_write.sum"])
                out.write(f"  {k[1]:34s} grid {k[2]:14s} {tms:8.3f} ms  {fl/1e9:7.2f} Gflop ({fl/tms/1e9:5.2f} TF/s)  dram rd {rd/1e9:6.3f} wr {wr/1e9:6.3f} GB "
                          f"({(rd+wr)/tms/1e6:6.0f} GB/s)  fp64 pipe {m['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'][0]:4.1f}%  "
                          f"inst {m['smsp__inst_executed.sum'][0]/1e6:7.1f} M\n")
                kind = "linearize" if k[1].startswith("linearize") else ("backward" if k[1].startswith("backward") else None)
                if kind:
                    acc[kind].append((tms, fl, rd + wr))
            if acc["linearize"]:
                passes = len(acc["linearize"]) / 3.0   # terminal-constraint workloads: knots [0, T-2), T-2, T-1 are three launches
                counts[label]["linearize"] = {"fp64_flops_per_knot": sum(a[1] for a in acc["linearize"]) / (passes * B * T),
                                              "dram_bytes_per_knot": sum(a[2] for a in acc["linearize"]) / (passes * B * T),
                                              "ms_per_pass": sum(a[0] for a in acc["linearize"]) / passes}
            if acc["backward"]:
                a = min(acc["backward"])
                counts[label]["backward"] = {"fp64_flops_per_knot": a[1] / (B * T), "dram_bytes_per_knot": a[2] / (B * T), "ms_clean_pass": a[0]}
    json.dump(counts, open(os.path.join(P, "r02_kernel_counts.json"), "w"), indent=1)
    with open(os.path.join(P, "r02_launch_shares.txt"), "w") as out:
        for tag in ("onestream", "default"):
            f = os.path.join(G, f"r02_launches_bench_c2_{tag}.csv")
            if not os.path.exists(f):
                continue
            shutil.copy(f, os.path.join(P, f"r02_launches_bench_c2_{tag}.csv"))
            rows = list(csv.reader(open(f)))
            hdr = None; agg = collections.OrderedDict(); tot = 0.0
            for r in rows:
                if "Kernel Name" in r:
                    hdr = r; continue
                if hdr and len(r) == len(hdr):
                    d = dict(zip(hdr, r))
                    if d.get("Metric Name") == "gpu__time_duration.sum":
                        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
                        v = v / 1e3 if u == "ns" else (v if u in ("us", "usecond") else v * 1e3)
                        agg.setdefault(d["Kernel Name"].split("(")[0].replace("void ", "").replace("ddpb::", ""), []).append(v); tot += v
            out.write(f"== python bench.py --steps 2 --warmup 3 --no-cpu-baseline under ncu --metrics gpu__time_duration.sum --clock-control none "
                      f"({'DDPB_SUBBATCHES=1' if tag == 'onestream' else 'default sub-batches, ncu serialises their kernels'})\n")
            out.write(f"   {sum(len(v) for v in agg.values())} launches captured, {tot/1e3:.1f} ms of kernel time\n")
            for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
                out.write(f"   {k:34s} n={len(v):4d}  sum {sum(v)/1e3:8.2f} ms  share {100*sum(v)/tot:5.1f}%  avg {sum(v)/len(v):8.1f} us\n")
    print("profiles written")


if __name__ == "__main__":
    main()
