#!/usr/bin/env python3
"""Where the two codec kernels spend their cycles, phase by phase (profiling build with clock64() marks:
`python -m yuv-manipulations-2_b200.build --clk`, loaded through MYYUVB_LIB_VARIANT=clk).

  python profiles/phase_clocks.py [frames] > profiles/r02_phase_clocks.json

Per workload: the share of warp-resident clocks between consecutive marks (lane 0 of every warp, summed over all warps) and
the launch sequence times of the same build.  Shares, not absolute times: the marks cost a few instructions each."""
import importlib
import json
import os
import pathlib
import sys

os.environ["MYYUVB_LIB_VARIANT"] = "clk"
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module("yuv-manipulations-2_b200")
synth = importlib.import_module("yuv-manipulations-2_b200.synth")
sys.argv = sys.argv[:1] + sys.argv[1:]
ENC = ["ticket", "load+dct+quant", "block sort", "histogram", "deferral queue", "plan (order/heap/merges/sizes)", "size scan+reserve",
       "emit (sort/codes/table/stream)", "barrier after emit", "copy out+barrier"]
DEC = ["ticket+barrier", "stage+scan+zero", "block sort", "entropy decode", "idct+store"]
W, H = 3840, 2160


def natural(n):
    import struct

    blob = (ROOT / "oracle/_ref/golden/chef-with-trumpet.myyuv").read_bytes()
    _, _, _, _, _, _, w, h, pos = struct.unpack_from("<2sIIHIIIII", blob, 0)
    base = np.frombuffer(blob, np.uint8)[pos: pos + w * h * 3 // 2].copy()
    return synth.tiled_real_iyuv(base, w, h, W, H, n, 0)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda", 0)
    ctx = pkg.Context(0)
    out = {"frames": n, "workloads": []}
    for name, q, frames in (("noise-grad", 50, None), ("tiled-real", 50, "nat"), ("tiled-real", 90, "nat"), ("noise-grad", 90, None)):
        d_in = synth.iyuv_frames_torch(W, H, n, dev) if frames is None else torch.from_numpy(natural(n)).to(dev)
        cap = n * 20 * 1024 * 1024
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        d_back = torch.empty_like(d_in)
        qq = (q,) * 3
        torch.cuda.synchronize()  # the tensors above were filled on torch's stream, the library runs on its own
        for _ in range(2):
            ctx.compress_batch_dev(d_in, W, H, qq, n, d_out, cap, d_off)
            ctx.decompress_batch_dev(d_out, d_off, W, H, qq, n, d_back)
        ctx.batch_status()
        pkg.capi.phase_clocks(True)
        ctx.compress_batch_dev(d_in, W, H, qq, n, d_out, cap, d_off)
        cms = ctx.last_kernel_ms()
        ctx.decompress_batch_dev(d_out, d_off, W, H, qq, n, d_back)
        dms = ctx.last_kernel_ms()
        ctx.batch_status()
        clk = pkg.capi.phase_clocks(True).astype(np.float64)
        enc = {k: round(float(v / clk[0].sum()), 4) for k, v in zip(ENC, clk[0])}
        dec = {k: round(float(v / clk[1].sum()), 4) for k, v in zip(DEC, clk[1])}
        out["workloads"].append({"content": name, "quality": q, "compress_ms": round(cms, 4), "decompress_ms": round(dms, 4),
                                 "compress_phase_share": enc, "decompress_phase_share": dec})
        print(name, q, cms, dms, file=sys.stderr)
        del d_in, d_out, d_back
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
