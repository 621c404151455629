"""Turns the ncu artefacts a gpurun session leaves in gpurun_out/ (scripts/gpu_profiles.sh) into the small
tracked summaries under profiles/: the launch list of the bench command, per-kernel shares of the last
full-size commit in it, and the selected metrics of the full captures."""
import csv
import io
import json
import shutil
import subprocess
import sys

ROUND = sys.argv[1] if len(sys.argv) > 1 else "r1"
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.avg.per_second",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__sass_inst_executed_op_local_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
]


def capture(rep, source):
    import os
    raw = rep.replace(".ncu-rep", ".raw.csv")
    if os.path.exists(raw):  # exported on the GPU box (scripts/gpu_profiles.sh) when the report is too large to bring back
        out = open(raw).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    launches = []
    for v in vals:
        d = {"kernel": v[hdr.index("Kernel Name")][:60]}
        for w in WANT:
            if w in hdr:
                d[w] = (v[hdr.index(w)] + " " + units[hdr.index(w)]).strip()
        launches.append(d)
    return {"source": source, "launches": launches}


def launch_shares(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit() and r[12] == "gpu__time_duration.sum"]
    # a commit starts with its recode: msm_recode_count (partitioned recode, round 2) or msm_recode (plain)
    marker = "msm_recode_count" if any("msm_recode_count" in r[4] for r in rows) else "msm_recode"
    recodes = [i for i, r in enumerate(rows) if marker in r[4]]
    grid = lambda r: int(r[8].strip("()").split(",")[0])
    big = max(grid(rows[i]) for i in recodes)
    last = max(i for i in recodes if grid(rows[i]) == big)  # the last full-size commit of the run
    nxt = min([i for i in recodes if i > last] + [len(rows)])
    ends = [i for i in range(last, nxt) if "xyzz_to_affine_bytes" in rows[i][4]]  # a commit ends with its to-affine
    if ends:
        nxt = ends[0] + 1
    agg = {}
    for r in rows[last:nxt]:
        name = r[4].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / (1e6 if r[13] in ("ns", "nsecond") else 1e3 if r[13] in ("us", "usecond") else 1.0)
    total = sum(a[1] for a in agg.values())
    return {"source": "ncu --metrics gpu__time_duration.sum --clock-control none -c 800 python bench.py --steps 2 --warmup 3 --no-verify "
                      "--no-extras (profiles/launches_%s_bench_2p24.csv); the last full-size (2^24) commit of the run; serialised "
                      "cold-cache times - compare SHARES with bench.py's live phases_ms" % ROUND,
            "step_total_ms": round(total, 3),
            "kernels": [{"kernel": k, "launches": a[0], "ms": round(a[1], 4), "share": round(a[1] / total, 4)} for k, a in agg.items()]}


if __name__ == "__main__":
    if ROUND == "r1":
        raise SystemExit("round-1 summaries are frozen; run with r2")
    shutil.copy(f"gpurun_out/launches_{ROUND}_bench_2p24.csv", f"profiles/launches_{ROUND}_bench_2p24.csv")
    json.dump(launch_shares(f"profiles/launches_{ROUND}_bench_2p24.csv"), open(f"profiles/launch_shares_{ROUND}.json", "w"), indent=1)
    acc = capture(f"gpurun_out/prof_accumulate_{ROUND}.ncu-rep",
                  "ncu --set full --clock-control none --import-source on -k regex:msm_accumulate (scripts/one_commit.py 24, 2^24 points, c=22)")
    json.dump(acc, open(f"profiles/ncu_msm_accumulate_{ROUND}.json", "w"), indent=1)
    sc = capture(f"gpurun_out/prof_sort_{ROUND}.ncu-rep",
                 "ncu --set full --clock-control none --import-source on -k regex:sort_tile_scatter|sort_tile_hist|sort_group_local|msm_recode "
                 "(2^24 points, c=22, 201.3M pairs: partitioned recode, one 256-way pass inside the partitions, group-local sort)")
    json.dump(sc, open(f"profiles/ncu_sort_{ROUND}.json", "w"), indent=1)
    rd = capture(f"gpurun_out/prof_reduce_{ROUND}.ncu-rep",
                 "ncu --set full --clock-control none --import-source on -k regex:msm_bucket_chunks|msm_bucket_reduce|msm_merge_level|xyzz_tree_reduce (2^24, c=22)")
    json.dump(rd, open(f"profiles/ncu_merge_reduce_{ROUND}.json", "w"), indent=1)
    a = acc["launches"][0]
    gb = lambda s: float(s.split()[0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[s.split()[1]]
    traffic = gb(a["dram__bytes_read.sum"]) + gb(a["dram__bytes_write.sum"])
    try:
        t = json.load(open(f"profiles/accumulate_traffic_{ROUND}.json"))
    except Exception:
        t = {"source": "dram__bytes_read.sum + dram__bytes_write.sum of msm_accumulate per launch, ncu --set full (scripts/gpu_profiles.sh)"}
    t["2^24_n1"] = traffic
    json.dump(t, open(f"profiles/accumulate_traffic_{ROUND}.json", "w"), indent=1)
    print("traffic", traffic, "shares total", json.load(open(f"profiles/launch_shares_{ROUND}.json"))["step_total_ms"])
