"""Turn the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/.
    python scripts/summarize_profiles.py r01
Every part runs only if its raw input is present (scripts/profile.sh can run a subset of its steps);
profiles/spmv_traffic.json is updated key by key.
"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
TRAFFIC = f"{P}/spmv_traffic.json"
traffic = json.load(open(TRAFFIC)) if os.path.exists(TRAFFIC) else {}

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_page(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    return rr[0], rr[1], rr[2:]


def write_full(h, u, r, f):
    f.write(f"kernel: {r[h.index('Kernel Name')]}\n")
    for k in KEEP:
        if k in h: f.write(f"{k:75s} {r[h.index(k)]:>18s} {u[h.index(k)]}\n")
    stalls = sorted(((float(r[i].replace(",", "") or 0), h[i]) for i in range(len(h))
                     if "issue_stalled" in h[i] and h[i].endswith("per_issue_active.ratio")), reverse=True)
    f.write("# warp stall reasons (warps per issue-active cycle)\n")
    for v, k in stalls[:8]: f.write(f"{k:75s} {v:18.3f}\n")


def dram_bytes(h, u, r):
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[h.index(name)].replace(",", "")) * UNIT[u[h.index(name)]]
    return tot


# 1) launch list -> per-kernel share of the step (setup kernels listed separately)
SETUP = ("cub::", "k_spgemm", "k_pick", "k_match", "k_leftover", "k_is_leader", "k_agg_id", "k_compose", "k_coarse_keys",
         "k_prolongator", "k_transpose", "k_split_keys", "k_rowptr", "k_make_keys", "k_fill_pattern", "k_head_flags",
         "k_assemble", "k_element", "k_inc_", "k_check_tris", "k_tile_nnz", "k_diag_inv", "k_dense_from", "k_dense_add",
         "k_dense_invert", "k_rowsum", "k_flag", "k_visc_vals", "k_inner_trig", "k_elem_thirds", "k_node_sum", "k_iota",
         "k_ptr_from", "k_fold_coo", "k_sell_", "k_to_f32", "k_gj_", "k_maxabs_bits", "k_sigma_keys")
if os.path.exists(f"{G}/{tag}_launches.csv"):
    rows = [r for r in csv.reader(open(f"{G}/{tag}_launches.csv")) if len(r) > 5]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += v
    is_setup = lambda k: any(t in k for t in SETUP)
    tot_step = sum(a[1] for k, a in agg.items() if not is_setup(k))
    tot_setup = sum(a[1] for k, a in agg.items() if is_setup(k))
    with open(f"{P}/{tag}_launch_list_summary.txt", "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 python bench.py --steps 1 --warmup 1 --no-cpu --no-extra\n")
        f.write("# (default configuration: AMG-preconditioned pressure CG; the V-cycle's graph nodes appear as kernels)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# launches captured: {sum(a[0] for a in agg.values())} (one-time setup + the warm-up and the timed step); "
                f"step kernels {tot_step/1e6:.1f} ms, one-time setup kernels {tot_setup/1e6:.1f} ms\n")
        f.write("# ---- kernels of the time step (share of step-kernel time)\n")
        for k, (n, t) in sorted(((k, a) for k, a in agg.items() if not is_setup(k)), key=lambda kv: -kv[1][1]):
            f.write(f"{100*t/tot_step:7.3f}%  launches={n:5d}  total_us={t/1e3:12.1f}  {k}\n")
        f.write("# ---- one-time setup (mesh topology, assembly, AMG hierarchy)\n")
        for k, (n, t) in sorted(((k, a) for k, a in agg.items() if is_setup(k)), key=lambda kv: -kv[1][1]):
            f.write(f"{100*t/tot_setup:7.3f}%  launches={n:5d}  total_us={t/1e3:12.1f}  {k[:110]}\n")
    os.system(f"gzip -c {G}/{tag}_launches.csv > {P}/{tag}_launches.csv.gz")
    print(open(f"{P}/{tag}_launch_list_summary.txt").read()[:2200])

# 2) full-set capture of the two big k_spmv_sell instances
if os.path.exists(f"{G}/{tag}_spmv_sell.ncu-rep"):
    h, u, rs = raw_page(f"{G}/{tag}_spmv_sell.ncu-rep")
    with open(f"{P}/{tag}_spmv_sell_ncu_full.txt", "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on --kernel-name-base demangled "
                "-k regex:'k_spmv_sell<.bool.(0|1), .bool.1' --launch-skip 6 -c 2 python scripts/prof_amg.py 6\n"
                "# (4M-triangle pressure operator, AMG-preconditioned CG; template arguments <SPLIT, DOT, F32>: <1,1,1> = finest\n"
                "#  up-sweep of the folded V-cycle, <0,1,0> = the CG's A*p; cold cache, one launch each)\n")
        for r in rs:
            write_full(h, u, r, f)
            name = r[h.index("Kernel Name")]
            key = "sell_up0_dram_bytes_per_launch" if "<1, 1, 1>" in name or "(bool)1, (bool)1, (bool)1" in name else "sell_ap_dram_bytes_per_launch"
            traffic[key] = dram_bytes(h, u, r)
            f.write("\n")
    traffic["sell_how"] = "ncu --set full capture of one launch of each k_spmv_sell instance (cold cache)"
    print(open(f"{P}/{tag}_spmv_sell_ncu_full.txt").read())

# 3) full-set capture of the persistent CG kernel
if os.path.exists(f"{G}/{tag}_cg_persistent.ncu-rep"):
    h, u, rs = raw_page(f"{G}/{tag}_cg_persistent.ncu-rep")
    with open(f"{P}/{tag}_cg_persistent_ncu_full.txt", "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:k_cg_persistent -c 1 python scripts/prof_cg.py 200\n")
        f.write("# (200 CG iterations of the 4M-triangle pressure operator in ONE launch; ncu flushes caches before the launch only)\n")
        write_full(h, u, rs[0], f)
    print(open(f"{P}/{tag}_cg_persistent_ncu_full.txt").read())

# 4) DRAM traffic of the persistent kernel with caches left alone
if os.path.exists(f"{G}/{tag}_cg_dram_nocachectl.csv"):
    iters = 200
    d = {}
    for row in csv.reader(open(f"{G}/{tag}_cg_dram_nocachectl.csv")):
        if len(row) > 14 and row[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_sector_hit_rate.pct"):
            d[row[12]] = float(row[14].replace(",", ""))
    traffic.update({"kernel": "k_cg_persistent", "iterations_in_launch": iters,
                    "dram_bytes_per_launch": d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"],
                    "dram_bytes_per_iteration": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / iters,
                    "dram_read_per_iteration": d["dram__bytes_read.sum"] / iters,
                    "dram_write_per_iteration": d["dram__bytes_write.sum"] / iters,
                    "us_per_iteration_under_ncu": d["gpu__time_duration.sum"] / iters / 1e3,
                    "lts_hit_rate_pct": d.get("lts__t_sector_hit_rate.pct"),
                    "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none (single pass, no replay, caches left alone)"})
    os.system(f"cp {G}/{tag}_cg_dram_nocachectl.csv {P}/{tag}_cg_dram_nocachectl.csv")

# 5) the rest of the path (assembly, div / grad, locator, tracers): DRAM bytes and duration per launch -> roofline fractions
if os.path.exists(f"{G}/{tag}_path_kernels.csv"):
    rows = [r for r in csv.reader(open(f"{G}/{tag}_path_kernels.csv")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, ui, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        if r[mi].startswith("dram__bytes"): v *= UNIT.get(r[ui], 1)
        if r[mi] == "gpu__time_duration.sum": v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        per.setdefault((r[idi], r[ki].split("(")[0]), {})[r[mi]] = v
    agg = collections.OrderedDict()
    for (_, k), d in per.items():
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0); a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        a[3] = d.get("launch__registers_per_thread", 0); a[4] += d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0)
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    # algorithmic bytes per launch on the 4M-triangle mesh (T = 4 194 304, N = 2 099 200, nnz = 14 690 826)
    T, N, NNZ, P4 = 4194304, 2099200, 14690826, 4000000
    ALG = {"k_element_stiffness": 12 * T + 16 * N + 72 * T, "k_assemble": 72 * T + 36 * T + 4 * NNZ + 8 * NNZ,
           "k_div_elem": 12 * T + 16 * N + 16 * N + 8 * T, "k_div_node": 4 * N + 12 * T + 8 * T + 8 * N + 8 * N,
           "k_grad_elem": 12 * T + 16 * N + 8 * N + 16 * T, "k_grad_node": 4 * N + 12 * T + 16 * T + 8 * N + 16 * N,
           "k_tracer_step": (16 + 16 + 4 + 4 + 4 + 4) * P4, "k_advect_dye": 16 * N + 16 * N + 8 * N + 8 * N, "k_locate_knn": 16 * N + 4 * N}
    with open(f"{P}/{tag}_path_kernels.txt", "w") as f:
        f.write("# ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,... --clock-control none python scripts/prof_path.py\n"
                "# 4M-triangle mesh (N = 2 099 200), 4M tracers; per launch, mean over the captured launches (cold cache, serialised)\n"
                f"# roofline: algorithmic bytes (DESIGN.md section 4) / duration against the measured copy bandwidth {peak:.1f} GB/s\n")
        f.write(f"{'kernel':40s} {'launches':>8s} {'us':>9s} {'DRAM MB':>9s} {'alg MB':>9s} {'alg GB/s':>9s} {'frac':>6s} {'regs':>5s} {'warps %':>8s}\n")
        for k, (n, t, b, regs, wa) in agg.items():
            base = k.split("<")[0].replace("void ", "").replace("fs::", "").strip()
            alg = next((v for kk, v in ALG.items() if base.startswith(kk)), None)
            us, mb = t / n, b / n / 1e6
            gbs = (alg / (us * 1e-6) / 1e9) if alg else float("nan")
            f.write(f"{k[:40]:40s} {n:8d} {us:9.1f} {mb:9.1f} {(alg or 0)/1e6:9.1f} {gbs:9.0f} {gbs/peak:6.2f} {int(regs):5d} {wa/n:8.1f}\n")
    print(open(f"{P}/{tag}_path_kernels.txt").read())

# 6) A/B of the bulk-async (TMA) staged SELL kernel against the register-staged one
if os.path.exists(f"{G}/{tag}_spmv_sell_bulk.ncu-rep"):
    h, u, rs = raw_page(f"{G}/{tag}_spmv_sell_bulk.ncu-rep")
    with open(f"{P}/{tag}_spmv_sell_bulk_ncu_full.txt", "w") as f:
        f.write("# FS_SELL_TMA=2 ncu --set full --clock-control none --import-source on --kernel-name-base demangled "
                "-k regex:'k_spmv_sell_bulk<.bool.(0|1), .bool.1' --launch-skip 6 -c 2 python scripts/prof_amg.py 6\n"
                "# the same two launches as in the k_spmv_sell capture next to this file, run by the cp.async.bulk + mbarrier staged\n"
                "# kernel (2 stages per warp, 6 CTAs per SM for fp32 / 4 for fp64): compare duration, DRAM throughput, stall reasons\n")
        for r in rs:
            write_full(h, u, r, f)
            f.write("\n")
    print(open(f"{P}/{tag}_spmv_sell_bulk_ncu_full.txt").read())


# 7) the five k_spmv_sell launches of one AMG-PCG iteration with the packed (fp16 | 16-bit offset) V-cycle operators
if os.path.exists(f"{G}/{tag}_spmv_pk.ncu-rep"):
    h, u, rs = raw_page(f"{G}/{tag}_spmv_pk.ncu-rep")
    with open(f"{P}/{tag}_spmv_pk_ncu_full.txt", "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_spmv_sell<' "
                "--launch-skip 9 -c 5 python scripts/prof_amg.py 6\n"
                "# (one AMG-PCG iteration at 4M triangles: A*p <0,1,0,0> fp64 values; finest restriction <0,0,4,0>, level-1 restriction\n"
                "#  <0,0,2,0>, level-1 up-sweep <1,0,2,0>, finest up-sweep <1,1,4,0>; template arguments <SPLIT, DOT, FMT, DIST>,\n"
                "#  FMT 2 = packed entries with fp64 gathers, 4 = packed entries with fp32 gathers; cold cache, one launch each)\n")
        for r in rs:
            write_full(h, u, r, f)
            f.write("\n")
            name = r[h.index("Kernel Name")]
            if "<1, 1, 4, 0>" in name: traffic["sell_up0_dram_bytes_per_launch"] = dram_bytes(h, u, r)
            if "<0, 1, 0, 0>" in name: traffic["sell_ap_dram_bytes_per_launch"] = dram_bytes(h, u, r)
    print(open(f"{P}/{tag}_spmv_pk_ncu_full.txt").read())

json.dump(traffic, open(TRAFFIC, "w"), indent=1)
print(open(TRAFFIC).read())
