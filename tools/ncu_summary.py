"""Summarises an .ncu-rep (ncu --set full) into a small JSON for profiles/: per launch duration, DRAM bytes, registers,
occupancy, instruction count, IPC and the main stall reasons.  Usage: python tools/ncu_summary.py in.ncu-rep out.json [pictures in the capture]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_per_sm",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
}


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]]}
        for k, name in KEYS.items():
            if k in idx:
                try:
                    v = float(r[idx[k]])
                except ValueError:
                    v = r[idx[k]]
                d[name] = v
                if name.endswith("_MB") and units[idx[k]].lower().startswith("kbyte"):
                    d[name] = v / 1e3
                if name.endswith("_MB") and units[idx[k]].lower() == "byte":
                    d[name] = v / 1e6
        res.append(d)
    doc = {"source": rep, "command": "ncu --profile-from-start off --set full --clock-control none --import-source on python tools/ncu_picture.py "
                                     "(one benchmark picture, L2 flushed before it; DESIGN.md section 6)", "launches": res}
    if len(sys.argv) > 3:
        doc["pictures"] = int(sys.argv[3])      # pictures inside the capture window (tools/ncu_picture.py: 1)
    json.dump(doc, open(out, "w"), indent=1)
    print("wrote", out, len(res), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
