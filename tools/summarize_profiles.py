"""Turns the raw ncu outputs of tools/make_profiles.sh (gpurun_out/) into the tracked summaries under profiles/ (dev tool).
usage: summarize_profiles.py [round_tag]"""
import csv, io, json, os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(R, "gpurun_out"); P = os.path.join(R, "profiles")
# ---- launch list
src = os.path.join(G, tag + "_launches_mhpc.csv")
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = {}
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum": continue
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    k = r["Kernel Name"].split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, tag + "_launches_mhpc_summary.csv"), "w") as f:
    f.write("# ncu launch list summary, round %s - `python tools/profile_cmd.py mhpc 4096` (one 4096-problem MHPC trot solve)\n" % tag)
    f.write("# source: tools/make_profiles.sh (ncu --metrics gpu__time_duration.sum --clock-control none); times are cold-cache, serialised\n")
    f.write("kernel,launches,total_ms,share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]): f.write("%s,%d,%.3f,%.3f\n" % (k, a[0], a[1], a[1] / tot))
with open(os.path.join(P, tag + "_launches_mhpc.csv"), "w") as f: f.write("".join(lines))
# ---- full captures
out = open(os.path.join(G, tag + "_full_mhpc_raw.csv")).read()
rr = list(csv.reader(io.StringIO(out))); hdr, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed.sum"]
traffic = {}
with open(os.path.join(P, tag + "_ncu_full_mhpc.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none `python tools/profile_cmd.py mhpc 4096 1 3` (tools/make_profiles.sh), round %s\n" % tag)
    for r in rr[2:]:
        name = r[hdr.index("Kernel Name")]; f.write("== %s  grid %s block %s\n" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        for h in want:
            if h in d: f.write("  %s = %s %s\n" % (h, d[h], u[h]))
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(d[h] or 0) > 0.3: f.write("  %s = %s\n" % (h, d[h]))
        # executed fp64 flops, counted by the hardware: thread-level DFMA (x2), DADD, DMUL + tensor-pipe fp64 ops, per elapsed cycle, against the
        # DFMA peak of the chip (sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained x 2 flop per cycle)
        try:
            g = lambda h: float(d[h].replace(",", ""))
            scal = 2 * g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") + g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") + g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")
            tens = g("sm__ops_path_tensor_src_fp64.sum.per_cycle_elapsed")
            peak = 2 * g("sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained")
            f.write("  executed_fp64_flop_per_cycle: scalar pipe %.1f + tensor pipe %.1f = %.1f of %.0f (%.3f of the fp64 peak)\n" % (scal, tens, scal + tens, peak, (scal + tens) / peak))
        except (KeyError, ValueError):
            pass
        key = name.split("(")[0].split("<")[0].replace("void ", "").replace("cafe_dev::", "")
        def gb(h): return float(d[h]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u[h]]
        if key not in traffic: traffic[key] = {"dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"), "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]]}
json.dump({"source": "profiles/%s_ncu_full_mhpc.txt (one launch with all 4096 problems active)" % tag, "kernels": traffic}, open(os.path.join(P, tag + "_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, tag + "_launches_mhpc_summary.csv")).read()); print(json.dumps(traffic, indent=1))
