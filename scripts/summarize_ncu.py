#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the tracked text summaries under profiles/.

  summarize_ncu.py TAG [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof_spmv.ncu-rep] [--name spmv]
                       [--cmd "python bench.py ..."]

  --launches CSV : `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list
                   -> profiles/TAG_launches.txt (share / count / sum / max per kernel)
  --rep REP      : `ncu --set full` capture -> profiles/TAG_<name>_full.txt (the metrics the roofline discussion
                   uses) and the per-launch DRAM traffic of every captured kernel -> profiles/traffic.json
"""
import argparse, collections, csv, json, os, re, subprocess

ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--launches", default=None)
ap.add_argument("--rep", default=None)
ap.add_argument("--name", default="spmv")
ap.add_argument("--cmd", default="python bench.py --steps 1 --warmup 1 --no-cpu-baseline")
args = ap.parse_args()
tag, out = args.tag, "profiles"
os.makedirs(out, exist_ok=True)


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(<[^(]*>)?)", name)
    return (m.group(1) if m else name)[:90]


if args.launches and os.path.exists(args.launches):
    lines = [l for l in open(args.launches) if not l.startswith("==")]
    agg = collections.OrderedDict(); n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum": continue
        t = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        t = t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] = max(a[2], t); n += 1
    tot = sum(a[1] for a in agg.values())
    with open(f"{out}/{tag}_launches.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  {args.cmd}\n")
        f.write(f"# (per-launch times are cold-cache and serialised: compare SHARES)  launches={n} total_us={tot:.1f}\n")
        f.write(f"{'share':>6} {'count':>6} {'sum_us':>10} {'max_us':>9}  kernel\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{a[1]/tot*100:5.1f}% {a[0]:6d} {a[1]:10.1f} {a[2]:9.1f}  {k}\n")
    print(open(f"{out}/{tag}_launches.txt").read())

if args.rep and os.path.exists(args.rep):
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(f"{out}/{tag}_{args.name}_full.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({os.path.basename(args.rep)}; {args.cmd})\n")
        for r in rows[2:]:
            f.write("----\n")
            for w in want:
                if w in idx: f.write(f"{w} = {short(r[idx[w]]) if w == 'Kernel Name' else r[idx[w]]} {units[idx[w]]}\n")
    print(open(f"{out}/{tag}_{args.name}_full.txt").read())

    def num(v, u):
        x = float(v.replace(",", "")); return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    tr = {}
    for r in rows[2:]:
        k = re.sub(r"<.*", "", short(r[idx["Kernel Name"]])).split("::")[-1]
        t = num(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + num(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        tr.setdefault(k, []).append(t)
    old = {}
    try: old = json.load(open(f"{out}/traffic.json"))
    except Exception: pass
    old.update({k: int(max(v)) for k, v in tr.items()})   # the largest launch = the top-level one
    src = old.get("_sources", {})
    if not isinstance(src, dict): src = {}
    for k in tr: src[k] = f"{out}/{tag}_{args.name}_full.txt"
    old["_sources"] = src
    old.pop("_source", None)
    json.dump(old, open(f"{out}/traffic.json", "w"), indent=1)
    print("wrote", f"{out}/traffic.json", old)
