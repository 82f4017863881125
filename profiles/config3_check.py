"""BASELINE configs[2]: full pipeline XRGB -> IYUV -> DCT-50 on a synthetic 3840x2160 batch of 256 frames, device resident
(8.5 GB in), timed, with frames 0, 100 and 255 checked byte for byte against the oracle's two-step result."""
import importlib, json, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
pkg = importlib.import_module("yuv-manipulations-2_b200"); synth = importlib.import_module("yuv-manipulations-2_b200.synth")
import oracle
W, H, F, q = 3840, 2160, 256, (50, 50, 50)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = pkg.Context(0, stream.cuda_stream)
d_bg = torch.empty((F, H, W, 4), dtype=torch.uint8, device=dev)
for f0 in range(0, F, 8):
    d_bg[f0:f0 + 8] = synth.bgrx_frames_torch(W, H, 8, dev, first=f0)
cap = F * 4 * 1024 * 1024
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
d_off = torch.zeros(F + 1, dtype=torch.int64, device=dev)
out = {}
for chunk in (8, 32):
    ctx.xrgb_compress_batch_dev(d_bg, W, H, True, q, F, d_out, cap, d_off, None, chunk)
    ctx.batch_status()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    ctx.xrgb_compress_batch_dev(d_bg, W, H, True, q, F, d_out, cap, d_off, None, chunk)
    ev[1].record(stream)
    torch.cuda.synchronize()
    ctx.batch_status()
    ms = ev[0].elapsed_time(ev[1])
    pay = int(d_off[F].item())
    out[f"chunk{chunk}"] = {"ms": round(ms, 2), "Mpixel_s": round(F * W * H / ms / 1e3, 1), "payload_bytes": pay,
                            "GBps_fused_algorithmic": round((F * W * H * 4 + pay) / ms / 1e6, 1)}
ora = oracle.Oracle()
off = d_off.cpu().numpy()
ok = True
for f in (0, 100, 255):
    bg = synth.bgrx_frames_numpy(W, H, 1, first=f)[0]
    want = ora.compress(ora.bgrx_to_iyuv(bg, W, H, True), W, H, q)
    got = d_out[int(off[f]): int(off[f + 1])].cpu().numpy()
    ok = ok and np.array_equal(got, want)
out["frames_0_100_255_match_oracle"] = bool(ok)
print(json.dumps(out))
