#!/usr/bin/env python3
"""BASELINE.json configs[0..2] on one B200 with the final kernels of the round, each with its parity gate and with the unmodified
reference (oracle/_ref, OpenMP build) timed beside it on the box's cores.  One JSON object on stdout (kept as
profiles/r02_configs.json).

  configs[0]  chef-with-trumpet.myyuv (992x736): compress DCT 50 -> bytes == the shipped chef-with-trumpet-DCT-50.myyuv payload,
              decompress -> sha256 of the reference's decode (tests/golden/golden.json)
  configs[1]  the 4032x3008 image (stand-in for the missing -big.myyuv: the reference's decode of the shipped -big-DCT-50 file,
              SURVEY 8(d) config 2): compress / decompress DCT 90, bytes == the reference's on the same input
  configs[2]  full pipeline XRGB8888 -> IYUV -> DCT-50 on a synthetic 3840x2160 batch of 256 frames, device resident;
              frames 0, 100, 255 byte for byte against the oracle
  colour      xrgb_to_iyuv_kernel on 32 4K frames against the measured HBM copy bandwidth

Device times: CUDA events inside the library (myyuvb_last_kernel_ms) for the codec launch sequences, torch events on the
library's stream for the pipeline and the colour conversion.  Host-pointer times: wall clock around the C ABI call (pageable
numpy buffers, what YUV::compress / YUV::decompress pay from the second call of a process on)."""
import hashlib, importlib, json, pathlib, statistics, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
pkg = importlib.import_module("yuv-manipulations-2_b200"); synth = importlib.import_module("yuv-manipulations-2_b200.synth")
import oracle

GOLD = ROOT / "oracle/_ref/golden"
golden = json.loads((ROOT / "tests/golden/golden.json").read_text())
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = pkg.Context(0, stream.cuda_stream)   # device-pointer calls, on torch's current stream
hctx = pkg.Context(0)                      # host-pointer calls
ora = oracle.Oracle()
ref = oracle.Reference("omp") if oracle.have_reference() else None
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6550.0) if (ROOT / "MEASURED_PEAKS.json").exists() else 6550.0
out = {"hbm_peak_GBps": peak, "reference_threads": ref.threads if ref else None}
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def wall(fn, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); ts.append(1e3 * (time.perf_counter() - t0))
    return r, round(min(ts), 3)


def ref_ms(fn, n=3):
    """the reference's own clock around its call (ref_shim.cpp), best of n"""
    ts = []
    for _ in range(n):
        r = fn(); ts.append(1e3 * ref.last_seconds)
    return r, round(min(ts), 3)


def single_image(name, iyuv, w, h, q, copies):
    """one image: host-pointer call (wall clock), `copies` device-resident copies of it (library events), reference beside"""
    qq = (q, q, q)
    hctx.compress(iyuv, w, h, qq)  # first call of the process sizes the buffers
    pay, c_ms = wall(lambda: hctx.compress(iyuv, w, h, qq))
    back, d_ms = wall(lambda: hctx.decompress(pay, w, h, qq))
    d_in = torch.from_numpy(iyuv).to(dev).repeat(copies, 1).contiguous()
    cap = copies * pkg.capi.compress_bound(w, h)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev); d_off = torch.zeros(copies + 1, dtype=torch.int64, device=dev)
    d_back = torch.empty_like(d_in)
    cs, ds = [], []
    for it in range(7):
        ctx.compress_batch_dev(d_in, w, h, qq, copies, d_out, cap, d_off); c1 = ctx.last_kernel_ms()
        ctx.decompress_batch_dev(d_out, d_off, w, h, qq, copies, d_back); d1 = ctx.last_kernel_ms()
        if it >= 2: cs.append(c1); ds.append(d1)
    ctx.batch_status()
    off = d_off.cpu().numpy()
    same_dev = all(np.array_equal(d_out[int(off[i]): int(off[i + 1])].cpu().numpy(), pay) for i in (0, copies - 1)) and \
        bool(torch.equal(d_back[copies - 1].cpu(), torch.from_numpy(back)))
    px = w * h
    alg = px * 3 // 2 + pay.size
    r = {"image": f"{w}x{h}", "quality": q, "payload_bytes": int(pay.size), "bytes_per_pixel": round(pay.size / px, 4),
         "host_pointer_call_ms": {"compress": c_ms, "decompress": d_ms},
         "device_resident": {"copies": copies, "compress_ms": round(statistics.median(cs), 4), "decompress_ms": round(statistics.median(ds), 4),
                             "compress_Mpixel_s": round(copies * px / statistics.median(cs) / 1e3, 1),
                             "decompress_Mpixel_s": round(copies * px / statistics.median(ds) / 1e3, 1),
                             "compress_GBps": round(copies * alg / statistics.median(cs) / 1e6, 1),
                             "decompress_GBps": round(copies * alg / statistics.median(ds) / 1e6, 1),
                             "compress_frac_of_hbm": round(copies * alg / statistics.median(cs) / 1e6 / peak, 4),
                             "decompress_frac_of_hbm": round(copies * alg / statistics.median(ds) / 1e6 / peak, 4),
                             "same_bytes_as_host_pointer_call": bool(same_dev)}}
    if ref is not None:
        rp, rc_ms = ref_ms(lambda: ref.compress(iyuv, w, h, qq), 3)
        rb, rd_ms = ref_ms(lambda: ref.decompress(rp, w, h, qq), 3)
        r["reference_omp_ms"] = {"compress": rc_ms, "decompress": rd_ms}
        r["payload_equals_reference"] = bool(np.array_equal(rp, pay))
        r["decoded_equals_reference"] = bool(np.array_equal(rb, back))
    r["payload_equals_oracle"] = bool(np.array_equal(ora.compress(iyuv, w, h, qq), pay))
    del d_in, d_out, d_back
    out[name] = r
    return pay, back


# ---- configs[0] ----
small = oracle.read_myyuv(GOLD / "chef-with-trumpet.myyuv")
pay, back = single_image("configs0_small_q50", small["data"], small["w"], small["h"], 50, 64)
shipped = oracle.read_myyuv(GOLD / "chef-with-trumpet-DCT-50.myyuv")
out["configs0_small_q50"]["payload_equals_shipped_DCT50_file"] = bool(np.array_equal(shipped["data"], pay))
out["configs0_small_q50"]["decoded_sha256"] = sha(back)
out["configs0_small_q50"]["decoded_equals_reference_decode"] = sha(back) == golden["chef"]["decoded:chef-with-trumpet-DCT-50.myyuv"]

# ---- configs[1] ----
big = oracle.read_myyuv(GOLD / "chef-with-trumpet-big-DCT-50.myyuv")
bw, bh = big["w"], big["h"]
big_iyuv = hctx.decompress(big["data"], bw, bh, tuple(int(x) for x in big["params"]))
out["configs1_big_input_sha256"] = sha(big_iyuv)
out["configs1_big_input_equals_reference_decode"] = sha(big_iyuv) == golden["chef"]["decoded:chef-with-trumpet-big-DCT-50.myyuv"]
single_image("configs1_big_q90", big_iyuv, bw, bh, 90, 16)
single_image("configs1_big_q50", big_iyuv, bw, bh, 50, 16)

# ---- colour conversion ----
W, H, F = 3840, 2160, 32
bg = synth.bgrx_frames_torch(W, H, 4, dev).repeat(F // 4, 1, 1, 1).contiguous()
yuv = torch.empty((F, W * H * 3 // 2), dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3): ctx.xrgb_to_iyuv_batch_dev(bg, W, H, True, F, yuv)
torch.cuda.synchronize(); ev[0].record(stream)
for _ in range(10): ctx.xrgb_to_iyuv_batch_dev(bg, W, H, True, F, yuv)
ev[1].record(stream); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
out["colour_xrgb_to_iyuv"] = {"frames": F, "ms": round(ms, 4), "Mpixel_s": round(F * W * H / ms / 1e3, 1), "GBps": round(F * W * H * 5.5 / ms / 1e6, 1),
                              "frac_of_hbm": round(F * W * H * 5.5 / ms / 1e6 / peak, 3),
                              "frame0_equals_oracle": bool(np.array_equal(yuv[0].cpu().numpy(), ora.bgrx_to_iyuv(bg[0].cpu().numpy().reshape(-1), W, H, True)))}
del bg, yuv

# ---- configs[2] ----
F = 256
d_bg = torch.empty((F, H, W, 4), dtype=torch.uint8, device=dev)
for f0 in range(0, F, 8):
    d_bg[f0:f0 + 8] = synth.bgrx_frames_torch(W, H, 8, dev, first=f0)
cap = F * 4 * 1024 * 1024
d_out = torch.empty(cap, dtype=torch.uint8, device=dev); d_off = torch.zeros(F + 1, dtype=torch.int64, device=dev)
q = (50, 50, 50)
cfg2 = {"frames": F, "input_bytes": F * W * H * 4}
for chunk in (8, 32, 64):
    ctx.xrgb_compress_batch_dev(d_bg, W, H, True, q, F, d_out, cap, d_off, None, chunk); ctx.batch_status()
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); ev[0].record(stream)
        ctx.xrgb_compress_batch_dev(d_bg, W, H, True, q, F, d_out, cap, d_off, None, chunk)
        ev[1].record(stream); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
    ctx.batch_status()
    ms = statistics.median(ts); p = int(d_off[F].item())
    cfg2[f"chunk{chunk}"] = {"ms": round(ms, 3), "Mpixel_s": round(F * W * H / ms / 1e3, 1), "payload_bytes": p,
                             "GBps_fused_algorithmic": round((F * W * H * 4 + p) / ms / 1e6, 1),
                             "frac_of_hbm_fused_algorithmic": round((F * W * H * 4 + p) / ms / 1e6 / peak, 4)}
off = d_off.cpu().numpy(); ok = True
t0 = time.perf_counter()
for f in (0, 100, 255):
    b = synth.bgrx_frames_numpy(W, H, 1, first=f)[0]
    ok = ok and np.array_equal(d_out[int(off[f]): int(off[f + 1])].cpu().numpy(), ora.compress(ora.bgrx_to_iyuv(b, W, H, True), W, H, q))
cfg2["frames_0_100_255_equal_oracle"] = bool(ok)
if ref is not None:  # the reference on one frame: its converter is single-threaded by construction (myyuv_yuv.cpp:108)
    b = synth.bgrx_frames_numpy(W, H, 1, first=0)[0]
    y, conv_ms = ref_ms(lambda: ref.bgrx_to_iyuv(b, W, H, True), 2)
    _, comp_ms = ref_ms(lambda: ref.compress(y, W, H, q), 2)
    cfg2["reference_omp_ms_per_frame"] = {"convert": conv_ms, "compress": comp_ms}
out["configs2_pipeline_256x4k_q50"] = cfg2
print(json.dumps(out, indent=1))
