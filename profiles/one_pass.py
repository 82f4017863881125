#!/usr/bin/env python3
"""Two passes (compress + decompress) of 8 4K frames per workload, for a launch list under ncu:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file raw.csv python profiles/one_pass.py ng:50 nat:50 ...
   python profiles/one_pass.py --parse raw.csv ng:50 nat:50 ...      (second pass of every workload, one row per kernel)"""
import csv, importlib, pathlib, struct, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import os
W, H, N = 3840, 2160, int(os.environ.get("ONE_PASS_FRAMES", "8"))
if sys.argv[1] == "--parse":
    rows = list(csv.reader(open(sys.argv[2])))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    Hd = rows[hi]
    ci, ck, cv, cu = Hd.index("ID"), Hd.index("Kernel Name"), Hd.index("Metric Value"), Hd.index("Metric Unit")
    seq = []
    for r in rows[hi + 1:]:
        if len(r) > cv and r[ci].isdigit():
            name = r[ck].split("(")[0].split("::")[-1].replace("void ", "")
            us = float(r[cv].replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[cu], 1.0)
            seq.append((name, us))
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("dct_compress_kernel")] + [len(seq)]
    for w, name in enumerate(sys.argv[3:]):
        a, b = starts[2 * w + 1], starts[2 * w + 2]
        tot = sum(u for _, u in seq[a:b])
        print(f"== {name}: second pass, {N} frames, {tot:.1f} us of kernels")
        for k, u in seq[a:b]:
            print(f"   {k:34s} {u:9.1f} us {100 * u / tot:5.1f} %")
    sys.exit(0)
import numpy as np, torch
pkg = importlib.import_module("yuv-manipulations-2_b200")
synth = importlib.import_module("yuv-manipulations-2_b200.synth")
def natural(n):
    blob = (ROOT / "oracle/_ref/golden/chef-with-trumpet.myyuv").read_bytes()
    _, _, _, _, _, _, w, h, pos = struct.unpack_from("<2sIIHIIIII", blob, 0)
    return synth.tiled_real_iyuv(np.frombuffer(blob, np.uint8)[pos: pos + w * h * 3 // 2].copy(), w, h, W, H, n, 0)
dev = torch.device("cuda", 0)
ctx = pkg.Context(0)
for name in sys.argv[1:]:
    content, q = name.split(":")
    d_in = synth.iyuv_frames_torch(W, H, N, dev) if content == "ng" else torch.from_numpy(natural(N)).to(dev)
    cap = N * 20 * 1024 * 1024
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev); d_off = torch.zeros(N + 1, dtype=torch.int64, device=dev); d_back = torch.empty_like(d_in)
    torch.cuda.synchronize()
    for _ in range(2):
        ctx.compress_batch_dev(d_in, W, H, (int(q),) * 3, N, d_out, cap, d_off)
        ctx.decompress_batch_dev(d_out, d_off, W, H, (int(q),) * 3, N, d_back)
        ctx.batch_status()
    del d_in, d_out, d_back
