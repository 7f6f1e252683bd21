"""Turn the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/.
    python scripts/summarize_profiles.py r01
"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)

# 1) launch list -> per-kernel share of the step
rows = [r for r in csv.reader(open(f"{G}/{tag}_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"{P}/{tag}_launch_list_summary.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 1 --warmup 1 --no-cpu\n")
    f.write(f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write(f"# launches captured: {sum(a[0] for a in agg.values())}, total {tot/1e6:.1f} ms\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{100*t/tot:7.3f}%  launches={n:4d}  total_us={t/1e3:12.1f}  {k}\n")
os.system(f"cp {G}/{tag}_launches.csv {P}/{tag}_launches.csv")

# 2) full-set capture of the persistent CG kernel
raw = subprocess.run(["ncu", "-i", f"{G}/{tag}_cg_persistent.ncu-rep", "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h, u, r = rr[0], rr[1], rr[2]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
stalls = sorted(((float(r[i].replace(",", "")), h[i]) for i in range(len(h))
                 if "issue_stalled" in h[i] and h[i].endswith("per_issue_active.ratio")), reverse=True)
with open(f"{P}/{tag}_cg_persistent_ncu_full.txt", "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:k_cg_persistent -c 1 python scripts/prof_cg.py 200\n")
    f.write("# (200 CG iterations of the 4M-triangle pressure operator in ONE launch; ncu flushes caches before the launch only)\n")
    for k in keep:
        if k in h: f.write(f"{k:75s} {r[h.index(k)]:>18s} {u[h.index(k)]}\n")
    f.write("# warp stall reasons (warps per issue-active cycle)\n")
    for v, k in stalls[:8]: f.write(f"{k:75s} {v:18.3f}\n")
iters = 200
d = {}
for row in csv.reader(open(f"{G}/{tag}_cg_dram_nocachectl.csv")):
    if len(row) > 14 and row[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_sector_hit_rate.pct"):
        d[row[12]] = float(row[14].replace(",", ""))
per = (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / iters
json.dump({"kernel": "k_cg_persistent", "iterations_in_launch": iters,
           "dram_bytes_per_launch": d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"],
           "dram_bytes_per_iteration": per, "dram_read_per_iteration": d["dram__bytes_read.sum"] / iters,
           "dram_write_per_iteration": d["dram__bytes_write.sum"] / iters,
           "us_per_iteration_under_ncu": d["gpu__time_duration.sum"] / iters / 1e3,
           "lts_hit_rate_pct": d.get("lts__t_sector_hit_rate.pct"),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none (single pass, no replay, caches left alone)"},
          open(f"{P}/spmv_traffic.json", "w"), indent=1)
os.system(f"cp {G}/{tag}_cg_dram_nocachectl.csv {P}/{tag}_cg_dram_nocachectl.csv")
print(open(f"{P}/{tag}_launch_list_summary.txt").read()[:1500])
print(open(f"{P}/{tag}_cg_persistent_ncu_full.txt").read())
print(open(f"{P}/spmv_traffic.json").read())
