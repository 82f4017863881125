#!/usr/bin/env python3
"""Side measurements for profiles/r01_notes.md (run on the GPU box): the colour-conversion kernel against the HBM
roofline, single-image latency through the host-pointer C ABI (what the class API / CLI path pays), quality sweep."""
import importlib, json, os, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

pkg = importlib.import_module("yuv-manipulations-2_b200")
synth = importlib.import_module("yuv-manipulations-2_b200.synth")
out = {}
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
t0 = time.perf_counter()
ctx = pkg.Context(0, stream.cuda_stream)
out["context_create_ms"] = round(1e3 * (time.perf_counter() - t0), 1)

# ---- colour conversion kernel: 32 frames of 3840x2160 XRGB (1.06 GB in, 0.4 GB out) ----
W, H, F = 3840, 2160, 32
bg = synth.bgrx_frames_torch(W, H, 4, dev)
bg = bg.repeat(F // 4, 1, 1, 1).contiguous()
yuv = torch.empty((F, W * H * 3 // 2), dtype=torch.uint8, device=dev)
for _ in range(3):
    ctx.xrgb_to_iyuv_batch_dev(bg, W, H, True, F, yuv)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record(stream)
for _ in range(10):
    ctx.xrgb_to_iyuv_batch_dev(bg, W, H, True, F, yuv)
ev[1].record(stream)
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
byts = F * W * H * 5.5
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6650.0) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
out["xrgb_to_iyuv"] = {"frames": F, "ms": round(ms, 4), "Mpixel_s": round(F * W * H / ms / 1e3, 1), "GBps": round(byts / ms / 1e6, 1),
                       "frac_of_hbm_peak": round(byts / ms / 1e6 / peak, 3), "algorithmic_bytes": int(byts)}

# ---- the same frames as 24-bit BMP rows (B,G,R triplets): 4.5 bytes per pixel ----
bg24 = bg[..., :3].contiguous()
yuv24 = torch.empty_like(yuv)
for _ in range(3):
    ctx.bgr24_to_iyuv_batch_dev(bg24, W, H, True, F, yuv24)
torch.cuda.synchronize()
ev[0].record(stream)
for _ in range(10):
    ctx.bgr24_to_iyuv_batch_dev(bg24, W, H, True, F, yuv24)
ev[1].record(stream)
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
byts = F * W * H * 4.5
out["bgr24_to_iyuv"] = {"frames": F, "ms": round(ms, 4), "Mpixel_s": round(F * W * H / ms / 1e3, 1), "GBps": round(byts / ms / 1e6, 1),
                        "frac_of_hbm_peak": round(byts / ms / 1e6 / peak, 3), "algorithmic_bytes": int(byts),
                        "same_planes_as_32bit": bool(torch.equal(yuv24, yuv))}
del bg24, yuv24

# ---- full pipeline XRGB -> IYUV -> DCT-50 (BASELINE configs[2]), device resident, by chunk size ----
capp = F * pkg.capi.compress_bound(W, H)
p_out = torch.empty(capp, dtype=torch.uint8, device=dev)
p_off = torch.zeros(F + 1, dtype=torch.int64, device=dev)
for chunk in (1, 2, 3, 4, 8, 16, 32):
    for _ in range(2):
        ctx.xrgb_compress_batch_dev(bg, W, H, True, (50, 50, 50), F, p_out, capp, p_off, None, chunk)
    ctx.batch_status()
    torch.cuda.synchronize()
    ev[0].record(stream)
    for _ in range(5):
        ctx.xrgb_compress_batch_dev(bg, W, H, True, (50, 50, 50), F, p_out, capp, p_off, None, chunk)
    ev[1].record(stream)
    torch.cuda.synchronize()
    ctx.batch_status()
    ms = ev[0].elapsed_time(ev[1]) / 5
    pay = int(p_off[F].item())
    out[f"pipeline_chunk{chunk}"] = {"ms": round(ms, 3), "Mpixel_s": round(F * W * H / ms / 1e3, 1), "payload_bytes": pay,
                                    "GBps_fused_bytes": round((F * W * H * 4 + pay) / ms / 1e6, 1)}
del p_out

if os.environ.get("EXTRAS_ONLY") == "pipeline":
    print(json.dumps(out, indent=1))
    sys.exit(0)

# ---- single image latency through the host-pointer C ABI (pageable numpy buffers, like the class API) ----
hctx = pkg.Context(0)
for (w, h) in ((992, 736), (3840, 2160), (4032, 3008), (7680, 4320)):
    f = synth.iyuv_frames_numpy(w, h, 1)[0]
    q = (50, 50, 50)
    t0 = time.perf_counter(); p = hctx.compress(f, w, h, q); first = time.perf_counter() - t0
    ts, td = [], []
    for _ in range(5):
        t0 = time.perf_counter(); p = hctx.compress(f, w, h, q); ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); d = hctx.decompress(p, w, h, q); td.append(time.perf_counter() - t0)
    out[f"host_api_{w}x{h}"] = {"compress_first_call_ms": round(1e3 * first, 2), "compress_ms": round(1e3 * min(ts), 2),
                                "decompress_ms": round(1e3 * min(td), 2), "payload_bytes": int(p.size)}

# ---- quality sweep, device resident, 32 frames 4K (BASELINE configs[4]) ----
d_in = synth.iyuv_frames_torch(W, H, F, dev)
cap = F * pkg.capi.compress_bound(W, H)
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
d_off = torch.zeros(F + 1, dtype=torch.int64, device=dev)
d_back = torch.empty_like(d_in)
for q in (10, 50, 90, 100):
    qq = (q, q, q)
    for _ in range(2):
        ctx.compress_batch_dev(d_in, W, H, qq, F, d_out, cap, d_off)
        ctx.decompress_batch_dev(d_out, d_off, W, H, qq, F, d_back)
    ctx.batch_status()
    cs, ds = [], []
    for _ in range(5):
        ctx.compress_batch_dev(d_in, W, H, qq, F, d_out, cap, d_off); cs.append(ctx.last_kernel_ms())
        ctx.decompress_batch_dev(d_out, d_off, W, H, qq, F, d_back); ds.append(ctx.last_kernel_ms())
    ctx.batch_status()
    pay = int(d_off[F].item())
    out[f"q{q}"] = {"compress_ms": round(min(cs), 3), "decompress_ms": round(min(ds), 3), "bytes_per_pixel": round(pay / (F * W * H), 4),
                    "compress_Mpixel_s": round(F * W * H / min(cs) / 1e3, 1), "decompress_Mpixel_s": round(F * W * H / min(ds) / 1e3, 1)}
# ---- natural content (SURVEY 8(d)(i) "tiled-real": the reference's chef-with-trumpet.myyuv tiled to 4K), 16 frames ----
gold = ROOT / "oracle" / "_ref" / "golden" / "chef-with-trumpet.myyuv"
if gold.exists():
    raw = gold.read_bytes()
    base = np.frombuffer(raw, np.uint8, 992 * 736 * 3 // 2, 64).copy()
    FR = 16
    host = synth.tiled_real_iyuv(base, 992, 736, W, H, FR)
    r_in = torch.from_numpy(host).to(dev)
    r_out = torch.empty(FR * pkg.capi.compress_bound(W, H), dtype=torch.uint8, device=dev)
    r_off = torch.zeros(FR + 1, dtype=torch.int64, device=dev)
    r_back = torch.empty_like(r_in)
    for q in (50, 90):
        qq = (q, q, q)
        cs, ds = [], []
        for it in range(6):
            ctx.compress_batch_dev(r_in, W, H, qq, FR, r_out, r_out.numel(), r_off); c_ms = ctx.last_kernel_ms()
            ctx.decompress_batch_dev(r_out, r_off, W, H, qq, FR, r_back); d_ms = ctx.last_kernel_ms()
            if it:
                cs.append(c_ms); ds.append(d_ms)
        ctx.batch_status()
        pay = int(r_off[FR].item())
        out[f"tiled_real_q{q}"] = {"frames": FR, "compress_ms": round(min(cs), 3), "decompress_ms": round(min(ds), 3),
                                   "bytes_per_pixel": round(pay / (FR * W * H), 4),
                                   "compress_Mpixel_s": round(FR * W * H / min(cs) / 1e3, 1), "decompress_Mpixel_s": round(FR * W * H / min(ds) / 1e3, 1)}
print(json.dumps(out, indent=1))
