#!/usr/bin/env python3
"""Turn an `ncu --set full` report into the small JSON summary committed under profiles/ (the .ncu-rep files
themselves stay in gpurun_out/, which is scratch).  Usage: summarize_ncu.py report.ncu-rep > summary.json"""
import csv, io, json, subprocess, sys

KEYS = {
    "duration_us": "gpu__time_duration.sum",
    "warp_instructions": "smsp__inst_executed.sum",
    "threads_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "pipe_alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "pipe_fma_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "pipe_xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "icache_hit_pct": "sm__icc_request_hit_rate.pct",
    "registers_per_thread": "launch__registers_per_thread",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "shared_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        k = {"kernel": d["Kernel Name"].split("(")[0], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for name, metric in KEYS.items():
            v = d.get(metric)
            try:
                v = float(v.replace(",", ""))
            except (AttributeError, ValueError):
                v = None
            if v is not None and name.endswith("_bytes") and u.get(metric, "").lower().startswith("mbyte"):
                v *= 1e6
            if v is not None and name.endswith("_bytes") and u.get(metric, "").lower().startswith("gbyte"):
                v *= 1e9
            if v is not None and name.endswith("_bytes") and u.get(metric, "").lower().startswith("kbyte"):
                v *= 1e3
            if v is not None and name == "duration_us" and u.get(metric, "").lower().startswith("ns"):
                v /= 1e3
            if v is not None and name == "duration_us" and u.get(metric, "").lower().startswith("ms"):
                v *= 1e3
            k[name] = v
        stalls = {}
        tot = 0.0
        for key, v in d.items():
            if key.startswith("smsp__pcsamp_warps_issue_stalled_") and not key.endswith("_not_issued"):
                try:
                    stalls[key[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(v)
                    tot += float(v)
                except ValueError:
                    pass
        k["stall_samples_pct"] = {a: round(100 * b / tot, 1) for a, b in sorted(stalls.items(), key=lambda x: -x[1]) if tot and b / tot > 0.01}
        out.append(k)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
