import importlib, sys, time, json, os
sys.path.insert(0, '.')
import numpy as np
pkg = importlib.import_module("yuv-manipulations-2_b200"); synth = importlib.import_module("yuv-manipulations-2_b200.synth")
out = {}
hctx = pkg.Context(0)
for (w, h) in ((3840, 2160), (4032, 3008), (7680, 4320)):
    f = synth.iyuv_frames_numpy(w, h, 1)[0]
    q = (50, 50, 50)
    p = hctx.compress(f, w, h, q)
    ts, td = [], []
    for _ in range(9):
        t0 = time.perf_counter(); p = hctx.compress(f, w, h, q); ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); d = hctx.decompress(p, w, h, q); td.append(time.perf_counter() - t0)
    out[f"{w}x{h}"] = (round(1e3 * min(ts), 2), round(1e3 * min(td), 2), round(1e3 * sorted(ts)[4], 2), round(1e3 * sorted(td)[4], 2))
print(os.environ.get("MYYUVB_STAGING", "ring"), json.dumps(out))
