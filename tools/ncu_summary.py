"""profiles/*.json from an .ncu-rep: the raw-page metrics the roofline argument needs + the pc-sampling stall mix.
usage: python tools/ncu_summary.py gpurun_out/full_<tag>.ncu-rep profiles/<name>.json ["what this capture is"]"""
import collections
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__icc_request_hit_rate.pct", "gcc__average_cache_request_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    what = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    res = {"kernel": vals[col["Kernel Name"]], "grid": vals[col["Grid Size"]], "block": vals[col["Block Size"]], "what": what}
    for k in KEEP:
        if k in col:
            try:
                res[k] = {"value": float(vals[col[k]].replace(",", "")), "unit": units[col[k]]}
            except ValueError:
                pass
    rd, wr = res.get("dram__bytes_read.sum"), res.get("dram__bytes_write.sum")
    if rd and wr:
        res["traffic_bytes"] = rd["value"] * UNIT.get(rd["unit"], 1.0) + wr["value"] * UNIT.get(wr["unit"], 1.0)
    # pc sampling: share of all warp samples per stall reason
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    h = next((r for r in src if "stall_long_sb" in r), None)
    if h:
        ix = {n: i for i, n in enumerate(h)}
        tot = collections.Counter()
        for r in src[src.index(h) + 1:]:
            if len(r) < len(h):
                continue
            for n, i in ix.items():
                if n.startswith("stall_") and "Not Issued" not in n:
                    try:
                        tot[n] += int(r[i] or 0)
                    except ValueError:
                        pass
        s = sum(tot.values()) or 1
        res["pc_sampling_share"] = {n: round(v / s, 4) for n, v in tot.most_common() if v}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: res[k] for k in ("kernel", "gpu__time_duration.sum", "traffic_bytes") if k in res}))


if __name__ == "__main__":
    main()
