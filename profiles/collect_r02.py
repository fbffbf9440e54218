#!/usr/bin/env python
"""gpurun_out/r02_* (written by profiles/run_r02_evidence.sh on a B200) -> the committed round-2 summaries:

  profiles/r02_bench.json              the bench line of the evidence run
  profiles/r02_workloads.json          one bench line per BASELINE.json config, each with its clocks record
  profiles/r02_launches.csv            ncu launch list of `bench.py --steps 2 --warmup 3`
  profiles/r02_launches_summary.json   per-kernel means of that list, share of the step
  profiles/r02_f1_launches.json        one frame per call: per-kernel times, test and train settings
  profiles/r02_rows_ncu_summary.json   `ncu --set full` key metrics of rows::fwd_kernel
  profiles/r02_rows_sass.txt           SASS excerpt of rows::fwd_kernel<2,true>: UTMALDG, UBLKCP, SYNCS, [R+UR] taps
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and r[0].isdigit()]
    return [(r[4], r[8], r[7], float(r[14]) / (1000.0 if r[13] in ("ns", "nsecond") else 1.0)) for r in rows]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("b2d::", "")
    return name.split("(")[0][:80]


def frames_of(grid):
    g = [int(x) for x in re.findall(r"\d+", grid)]
    return g[1] if len(g) > 1 else 1


def main():
    out = {}
    # --- bench line
    line = json.loads(open(os.path.join(G, "r02_bench.json")).read().strip().splitlines()[-1])
    json.dump(line, open(os.path.join(P, "r02_bench.json"), "w"), indent=1)
    # --- workloads
    wl = [json.loads(l) for l in open(os.path.join(G, "r02_workloads.jsonl")) if l.strip()]
    json.dump({"command": "bash profiles/run_workloads.sh (python bench.py --workload W --steps 20 --warmup 3)",
               "lines": wl}, open(os.path.join(P, "r02_workloads.json"), "w"), indent=1)
    # --- launch list of the bench command
    src = os.path.join(G, "r02_launches.csv")
    shutil.copy(src, os.path.join(P, "r02_launches.csv"))
    L = launches(src)
    F = line["config"]["frames_per_step_per_gpu"]
    per = collections.OrderedDict()
    for name, grid, block, us in L:
        d = per.setdefault(short(name), {"launches": 0, "sum_us": 0.0, "launches_step": 0, "sum_us_step": 0.0})
        d["launches"] += 1
        d["sum_us"] += us
        if frames_of(grid) == F or (f"({F}," in grid.replace(" ", "")) or grid.replace(" ", "").startswith(f"({F},"):
            d["launches_step"] += 1
            d["sum_us_step"] += us
    summ = {}
    step_total = 0.0
    for k, d in per.items():
        if d["launches_step"]:
            mean = d["sum_us_step"] / d["launches_step"]
            summ[k] = {"launches_with_%d_frames" % F: d["launches_step"], "mean_us": round(mean, 1)}
            step_total += mean
    for k in summ:
        summ[k]["share_of_step"] = round(summ[k]["mean_us"] / step_total, 4)
    json.dump({"command": "ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b2d|rows::' -c 600 python bench.py --steps 2 "
                          "--warmup 3 --no-cpu-baseline --no-gpu-baseline",
               "note": "cold-cache, serialised per-launch times; launches whose grid spans all %d frames of a step are the "
                       "timed path, the others belong to the e2e host-buffer leg, the F = 1 latency leg and the parity gate" % F,
               "step_sum_us": round(step_total, 1), "kernels": summ},
              open(os.path.join(P, "r02_launches_summary.json"), "w"), indent=1)
    # --- one frame per call
    f1 = {}
    for tag in ("f1", "f1_train"):
        path = os.path.join(G, f"r02_{tag}_launches.csv")
        if not os.path.exists(path):
            continue
        Lf = launches(path)
        # the last iteration: from the last select launch on (the first kernel of a call)
        first = lambda l: "select_fused" in l[0] or "score_hist" in l[0]
        idx = max(i for i, l in enumerate(Lf) if first(l))
        it = Lf[idx:]
        # RoIAlign launches of that iteration follow in the list only if the script ends with them: take one full cycle
        names = [short(l[0]) for l in Lf]
        cyc_start = max(i for i in range(idx) if first(Lf[i])) if any(first(l) for l in Lf[:idx]) else idx
        cyc = Lf[cyc_start:idx]
        f1[tag] = {"kernels": [{"kernel": short(n), "grid": g, "block": b, "us": round(us, 2)} for n, g, b, us in cyc],
                   "sum_us": round(sum(us for *_, us in cyc), 1)}
    json.dump({"command": "ncu --metrics gpu__time_duration.sum --clock-control none python profiles/f1_latency.py waymo_test 3 [TRAIN]",
               "note": "one proposal_layer + RoIAlign forward call on ONE Waymo frame (what the reference API issues); "
                       "cold-cache serialised kernel times of one iteration", **f1},
              open(os.path.join(P, "r02_f1_launches.json"), "w"), indent=1)
    # --- ncu --set full of the rows kernel
    rep = os.path.join(G, "r02_rows_full.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    keys = subprocess.run([sys.executable, os.path.join(P, "ncu_keys.py")], input=raw, capture_output=True, text=True).stdout
    d = {"kernel": "rows::fwd_kernel<2,true> (RoIAlign forward 7x7, sampling_ratio 2), round-2 final",
         "frames_in_launch": 64,
         "command": "ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 python profiles/rows_ab.py --one 64 2"}
    for l in keys.splitlines():
        if " = " in l:
            k, v = l.split(" = ", 1)
            d[k] = v.strip()
    json.dump(d, open(os.path.join(P, "r02_rows_ncu_summary.json"), "w"), indent=1)
    # --- SASS excerpt
    so = os.path.join(ROOT, "faster_rcnn_pytorch_multimodal_b200", "libb2dglue.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    fn = [f for f in sass.split("Function :") if "rows10fwd_kernelILi2ELb1" in f.split("\n")[0]][0]
    lines = fn.split("\n")
    pick = [l.strip() for l in lines if re.search(r"UTMALDG|UBLKCP|SYNCS|LDS[^\n]*\+UR|FFMA2|CREDUX|UTMAPF|ATOMS", l)]
    cnt = collections.Counter(re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l).group(1).split(".")[0]
                              for l in lines if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l) and re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l))
    with open(os.path.join(P, "r02_rows_sass.txt"), "w") as f:
        f.write("cuobjdump -sass libb2dglue.so, function %s\n" % lines[0].strip())
        f.write("opcode histogram: " + ", ".join(f"{k} {v}" for k, v in cnt.most_common(40)) + "\n\n")
        f.write("TMA / bulk / mbarrier instructions and the first uniform-addressed taps:\n")
        seen = collections.Counter()
        for l in pick:
            op = re.search(r"(UTMALDG|UBLKCP|SYNCS|LDS|FFMA2|CREDUX|UTMAPF|ATOMS)", l).group(1)
            seen[op] += 1
            if seen[op] <= (30 if op == "LDS" else 8):
                f.write("  " + l + "\n")
    print("wrote profiles/r02_*")


if __name__ == "__main__":
    main()
