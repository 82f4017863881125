#!/usr/bin/env python3
"""Condense a raw ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
    ncu --clock-control none -k regex:'dct_|heavy_|scan_|place_|finalize_|parse_|dec_|publish_|sm_copy|iyuv' ... python bench.py --steps 1 --warmup 1
into one row per kernel of the LAST device-resident step (the last compress sequence + the last decompress sequence that
ran back to back) -> profiles/rNN_launches_*.csv, and the codec kernels' DRAM bytes -> profiles/rNN_traffic.json.
usage: launch_list.py raw.csv out.csv [traffic.json]"""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hi]
col = {n: H.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launches = {}
for r in rows[hi + 1:]:
    if len(r) <= col["Metric Value"] or not r[col["ID"]].isdigit():
        continue
    k = launches.setdefault(int(r[col["ID"]]), {"kernel": r[col["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "")})
    v = float(r[col["Metric Value"]].replace(",", ""))
    u = r[col["Metric Unit"]]
    name = r[col["Metric Name"]]
    if name == "gpu__time_duration.sum":
        k["duration_us"] = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        k["dram_read_bytes" if "read" in name else "dram_write_bytes"] = int(v * scale)
seq = [launches[i] for i in sorted(launches)]
# device-resident steps: a compress sequence directly followed by a decompress sequence; the bench's 64-frame steps are the
# ones with the longest coding kernel (the e2e leg runs the same kernels on 2-frame chunks); take the last of those
steps = []
for i, k in enumerate(seq):
    if not k["kernel"].startswith("dct_compress_kernel"):
        continue
    j = i + 1
    while j < len(seq) and not seq[j]["kernel"].startswith(("dct_compress_kernel", "dct_decompress_kernel")):
        j += 1
    if j < len(seq) and seq[j]["kernel"].startswith("dct_decompress_kernel"):
        steps.append((i, j))
longest = max(seq[i].get("duration_us", 0) for i, _ in steps)
start, end = [st for st in steps if seq[st[0]].get("duration_us", 0) > 0.9 * longest][-1]
step = seq[start:end + 1]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "duration_us", "dram_read_bytes", "dram_write_bytes"])
    for k in step:
        w.writerow([k["kernel"], round(k.get("duration_us", 0), 1), k.get("dram_read_bytes", 0), k.get("dram_write_bytes", 0)])
tot = sum(k.get("duration_us", 0) for k in step)
for k in step:
    print(f'{k["kernel"]:28s} {k.get("duration_us", 0):9.1f} us {100 * k.get("duration_us", 0) / tot:5.1f} %  read {k.get("dram_read_bytes", 0):>12d}  write {k.get("dram_write_bytes", 0):>12d}')
if len(sys.argv) > 3:
    t = {k["kernel"]: {"dram_read": k.get("dram_read_bytes", 0), "dram_write": k.get("dram_write_bytes", 0), "us": round(k.get("duration_us", 0), 1)}
         for k in step if k["kernel"] in ("dct_compress_kernel", "dct_decompress_kernel")}  # (template arguments are stripped from the names)
    src = "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one step of bench.py (" + sys.argv[2] + ")"
    # the hash of the kernel sources the capture was taken from: bench.py reports the traffic only for the same sources
    import hashlib, pathlib
    h = hashlib.sha256()
    csrc = pathlib.Path(__file__).resolve().parent.parent / "yuv-manipulations-2_b200" / "csrc"
    for name in ("kernels.cu", "block_codec.cuh", "kernels.h", "dct_matrix.inc"):
        h.update((csrc / name).read_bytes())
    ncu = None
    if len(sys.argv) > 4:  # summarize_ncu.py output of an `ncu --set full` capture of the two codec kernels
        full = json.load(open(sys.argv[4]))
        ncu = {k["kernel"].replace("void ", "").split("<")[0]: {f: k.get(f) for f in ("duration_us", "warp_instructions", "threads_per_instruction",
               "issue_active_pct", "pipe_alu_pct", "pipe_fma_pct", "pipe_lsu_pct", "warps_active_pct", "icache_hit_pct", "dram_throughput_pct",
               "stall_samples_pct")} for k in full}
        ncu["workload"] = sys.argv[5] if len(sys.argv) > 5 else "8 frames"
    import importlib
    sys.path.insert(0, str(csrc.parent.parent))
    code = importlib.import_module("yuv-manipulations-2_b200.build").device_code_sha256()  # the SASS of the library that ran
    json.dump({"frames": 64, "source": src, "sources_sha256": h.hexdigest(), "device_code_sha256": code, "ncu": ncu, **t}, open(sys.argv[3], "w"), indent=1)
