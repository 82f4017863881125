#!/usr/bin/env python3
"""What a CLI-style process pays: the FIRST compress / decompress call of a process through the host-pointer C ABI
(pageable buffers), with the allocation trace of the library (MYYUVB_TRACE=1).  One fresh process per image size.

  python profiles/first_call.py > profiles/r02_first_call.json"""
import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
CHILD = r'''
import importlib, json, sys, time
sys.path.insert(0, sys.argv[1])
w, h = int(sys.argv[2]), int(sys.argv[3])
t0 = time.perf_counter()
pkg = importlib.import_module("yuv-manipulations-2_b200")
synth = importlib.import_module("yuv-manipulations-2_b200.synth")
f = synth.iyuv_frames_numpy(w, h, 1)[0]
t1 = time.perf_counter()
ctx = pkg.Context(0)
t2 = time.perf_counter()
p = ctx.compress(f, w, h, (50, 50, 50))
t3 = time.perf_counter()
d = ctx.decompress(p, w, h, (50, 50, 50))
t4 = time.perf_counter()
p2 = ctx.compress(f, w, h, (50, 50, 50))
t5 = time.perf_counter()
d2 = ctx.decompress(p, w, h, (50, 50, 50))
t6 = time.perf_counter()
print(json.dumps({"w": w, "h": h, "ctx_create_ms": 1e3 * (t2 - t1), "compress_first_ms": 1e3 * (t3 - t2), "decompress_first_ms": 1e3 * (t4 - t3),
                  "compress_second_ms": 1e3 * (t5 - t4), "decompress_second_ms": 1e3 * (t6 - t5), "payload": int(p.size)}))
'''
out = []
for w, h in ((992, 736), (3840, 2160), (7680, 4320)):
    r = subprocess.run([sys.executable, "-c", CHILD, str(ROOT), str(w), str(h)], env=dict(os.environ, MYYUVB_TRACE="1"),
                       capture_output=True, text=True, timeout=600)
    rec = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"w": w, "h": h, "error": r.stderr[-2000:]}
    rec["alloc_trace"] = [l for l in r.stderr.splitlines() if l.startswith("[myyuvb]")]
    out.append(rec)
print(json.dumps(out, indent=1))
