#!/usr/bin/env python3
"""Slim same-box A/B (see ab.py) with an output check: every variant must produce the same payload bytes and the same decoded
frames (sha256 of the whole payload and of three decoded frames per workload).  usage: ab2.py variantA variantB [...]  ("" = product)"""
import json, os, subprocess, sys
CHILD = r'''
import hashlib, importlib, json, os, pathlib, struct, sys, statistics
ROOT = pathlib.Path(sys.argv[1]); sys.path.insert(0, str(ROOT))
import numpy as np, torch
pkg = importlib.import_module("yuv-manipulations-2_b200"); synth = importlib.import_module("yuv-manipulations-2_b200.synth")
W, H = 3840, 2160
def natural(n):
    blob = (ROOT / "oracle/_ref/golden/chef-with-trumpet.myyuv").read_bytes()
    _, _, _, _, _, _, w, h, pos = struct.unpack_from("<2sIIHIIIII", blob, 0)
    return synth.tiled_real_iyuv(np.frombuffer(blob, np.uint8)[pos: pos + w * h * 3 // 2].copy(), w, h, W, H, n, 0)
dev = torch.device("cuda", 0); ctx = pkg.Context(0); out = {}
for name, n in (("ng:50", 64), ("nat:50", 16), ("nat:90", 16), ("ng:90", 16)):
    content, q = name.split(":"); qq = (int(q),) * 3
    d_in = synth.iyuv_frames_torch(W, H, n, dev) if content == "ng" else torch.from_numpy(natural(n)).to(dev)
    cap = n * 20 * 1024 * 1024
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev); d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev); d_back = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    c, d = [], []
    for it in range(7):
        ctx.compress_batch_dev(d_in, W, H, qq, n, d_out, cap, d_off); c1 = ctx.last_kernel_ms()
        ctx.decompress_batch_dev(d_out, d_off, W, H, qq, n, d_back); d1 = ctx.last_kernel_ms()
        if it >= 2: c.append(c1); d.append(d1)
    ctx.batch_status()
    h = hashlib.sha256(d_out[: int(d_off[n].item())].cpu().numpy().tobytes())
    for f in (0, n // 2, n - 1): h.update(d_back[f].cpu().numpy().tobytes())
    out[name] = [round(statistics.median(c), 4), round(statistics.median(d), 4), h.hexdigest()[:16]]
    del d_in, d_out, d_back
print(json.dumps(out))
'''
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = sys.argv[1:]
rounds = int(os.environ.get("AB_ROUNDS", "2"))
res = {v: [] for v in variants}
for rnd in range(rounds):
    for v in variants:
        r = subprocess.run([sys.executable, "-c", CHILD, root], env=dict(os.environ, MYYUVB_LIB_VARIANT=v), capture_output=True, text=True)
        res[v].append(json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-500:]})
hashes = {}
for v, runs in res.items():
    print(repr(v))
    for r in runs:
        print("   ", r)
        for k, val in r.items():
            if isinstance(val, list): hashes.setdefault(k, set()).add(val[2])
print("outputs identical across variants and rounds:", {k: len(s) == 1 for k, s in hashes.items()})
