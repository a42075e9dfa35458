"""Condenses `ncu --page raw --csv` exports (gpurun_out/prof_<round>_*_raw.csv) into profiles/: one small CSV per capture with the
metrics DESIGN.md quotes, and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic)."""
import csv, glob, json, os, sys
RND = sys.argv[1] if len(sys.argv) > 1 else "r3"      # capture round: gpurun_out/prof_<RND>_*_raw.csv -> profiles/<RND>_*
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "lts__t_bytes.sum"]
traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
traffic = json.load(open(traffic_path))
keymap = {"lik_cfg3": "cfg3:1", "lik_cfg2": "cfg2:1", "lik_cfg5": "cfg5:1", "lik_cfg3sh": "cfg3:sharded-kernel-one-rank"}
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_%s_*_raw.csv" % RND))):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print("skip", path); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = os.path.basename(path)[len("prof_%s_" % RND):-len("_raw.csv")]
    out = os.path.join(ROOT, "profiles", "%s_%s_ncu_full_summary.csv" % (RND, name))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for k in ("Kernel Name", "Block Size", "Grid Size"):
            if k in hdr:
                w.writerow([k, "", vals[hdr.index(k)]])
        for k in KEEP:
            if k in hdr:
                w.writerow([k, units[hdr.index(k)], vals[hdr.index(k)]])
    def get(k):
        v, u = float(vals[hdr.index(k)]), units[hdr.index(k)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    dur = vals[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")]
    print("%-12s %s  duration %s  dram read %.1f MB write %.1f MB  L2 hit %s%%" % (name, vals[hdr.index("Kernel Name")][:60], dur, rd / 1e6, wr / 1e6,
          vals[hdr.index("lts__t_sector_hit_rate.pct")]))
    if name in keymap:
        traffic[keymap[name]] = {"kernel": vals[hdr.index("Kernel Name")], "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                                 "source": "profiles/%s_%s_ncu_full_summary.csv" % (RND, name)}
    elif name.startswith("tick"):
        traffic["sampler-%s:1" % name] = {"kernel": vals[hdr.index("Kernel Name")], "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                                          "note": "one launch = up to 64 ticks (leapfrogs of all 64 chains)", "source": "profiles/%s_%s_ncu_full_summary.csv" % (RND, name)}
json.dump(traffic, open(traffic_path, "w"), indent=1)
