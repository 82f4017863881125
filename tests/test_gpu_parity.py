"""Parity of the CUDA path (through the C ABI) with the oracle, the reference's golden files and the
reference itself.  Integer/byte work: bit-exact everywhere (no tolerance)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SMALL = [(16, 16), (32, 16), (16, 32), (48, 80), (256, 128), (992, 736), (1296, 720)]
QS = [(50, 50, 50), (90, 90, 90), (10, 10, 10), (1, 1, 1), (100, 100, 100), (50, 51, 75), (97, 3, 64)]


def frames(synth, w, h, n=1, first=0):
    return synth.iyuv_frames_numpy(w, h, n, first)


@pytest.mark.parametrize("w,h", SMALL)
def test_colour_conversion_matches_oracle(ctx, ora, synth, w, h):
    bgrx = synth.bgrx_frames_numpy(w, h, 1, first=3)[0]
    for bottom_up in (True, False):
        got = ctx.xrgb_to_iyuv(bgrx, w, h, bottom_up)
        assert np.array_equal(got, ora.bgrx_to_iyuv(bgrx, w, h, bottom_up))


@pytest.mark.parametrize("w,h", [(2, 2), (12, 10), (20, 6), (36, 4), (1002, 14)])
def test_colour_conversion_narrow_widths(ctx, ora, synth, w, h):
    # widths that are not a multiple of 8 take the one-quad-per-thread kernel (the reference only needs even sizes)
    bgrx = synth.bgrx_frames_numpy(w, h, 1, first=5)[0]
    for bottom_up in (True, False):
        assert np.array_equal(ctx.xrgb_to_iyuv(bgrx, w, h, bottom_up), ora.bgrx_to_iyuv(bgrx, w, h, bottom_up))


@pytest.mark.parametrize("w,h", SMALL + [(4, 2), (12, 10), (20, 6), (36, 4), (1004, 14), (3840, 2160)])
def test_bgr24_conversion_matches_oracle(ctx, ora, synth, w, h):
    """24-bit BMP rows: 8-pixel kernel for widths that are a multiple of 8, quad kernel otherwise."""
    bgr = np.ascontiguousarray(synth.bgrx_frames_numpy(w, h, 1, first=7)[0].reshape(-1, 4)[:, :3]).reshape(-1)
    for bottom_up in (True, False):
        assert np.array_equal(ctx.bgr24_to_iyuv(bgr, w, h, bottom_up), ora.bgr24_to_iyuv(bgr, w, h, bottom_up))


def test_bgr24_random_bytes_and_batch(ctx, ora, ref, pkg):
    import torch

    rng = np.random.default_rng(2424)
    w, h, n = 200, 64, 3
    frames24 = rng.integers(0, 256, (n, w * h * 3), dtype=np.uint8)
    for f in frames24[:2]:
        assert np.array_equal(ctx.bgr24_to_iyuv(f, w, h, True), ref.bgr24_to_iyuv(f, w, h, True))
    d_in = torch.from_numpy(frames24).cuda()
    d_out = torch.empty(n * w * h * 3 // 2, dtype=torch.uint8, device="cuda")
    ctx.bgr24_to_iyuv_batch_dev(d_in, w, h, False, n, d_out)
    ctx.sync()
    got = d_out.cpu().numpy().reshape(n, -1)
    for i in range(n):
        assert np.array_equal(got[i], ora.bgr24_to_iyuv(frames24[i], w, h, False))
    with pytest.raises(pkg.MyyuvError):
        ctx.bgr24_to_iyuv(frames24[0][: 3 * 3 * 2], 3, 2)   # odd width


def test_colour_conversion_extremes(ctx, ora):
    # every (B,G,R) on a coarse lattice plus the pure-blue quad whose Cb sum wraps to 0
    vals = np.array([0, 1, 2, 15, 16, 17, 63, 64, 127, 128, 129, 191, 200, 253, 254, 255], np.uint8)
    b, g, r = np.meshgrid(vals, vals, vals, indexing="ij")
    px = np.stack([b.ravel(), g.ravel(), r.ravel(), np.zeros(b.size, np.uint8)], 1)  # 4096 pixels
    w, h = 64, 64
    img = px.reshape(h, w, 4)
    img = np.repeat(np.repeat(img, 2, 0), 2, 1)  # uniform 2x2 quads, 128 x 128
    got = ctx.xrgb_to_iyuv(img, 128, 128, True)
    assert np.array_equal(got, ora.bgrx_to_iyuv(img, 128, 128, True))
    blue = np.zeros((16, 16, 4), np.uint8)
    blue[..., 0] = 255
    out = ctx.xrgb_to_iyuv(blue, 16, 16, True)
    assert out[0] == 29 and out[256] == 0 and out[256 + 64] == 108


@pytest.mark.parametrize("w,h", SMALL)
@pytest.mark.parametrize("q", QS[:4])
def test_compress_matches_oracle(ctx, ora, synth, w, h, q):
    f = frames(synth, w, h)[0]
    got = ctx.compress(f, w, h, q)
    want = ora.compress(f, w, h, q)
    assert got.size == want.size
    assert np.array_equal(got, want)


@pytest.mark.parametrize("q", QS)
def test_compress_decompress_edge_cases(ctx, ora, synth, q):
    w, h = 128, 128
    f = synth.edge_case_iyuv(w, h)
    got = ctx.compress(f, w, h, q)
    want = ora.compress(f, w, h, q)
    assert np.array_equal(got, want)
    assert np.array_equal(ctx.decompress(want, w, h, q), ora.decompress(want, w, h, q))


def test_random_noise_many_symbols(ctx, ora):
    # uniform noise at q=100: ~60 distinct symbols per block -> local-memory scratch path, all rehash steps
    rng = np.random.default_rng(5)
    w, h = 256, 256
    f = rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8)
    for q in ((100, 100, 100), (95, 90, 85)):
        got = ctx.compress(f, w, h, q)
        want = ora.compress(f, w, h, q)
        assert np.array_equal(got, want)
        assert np.array_equal(ctx.decompress(got, w, h, q), ora.decompress(want, w, h, q))


@pytest.mark.parametrize("q", [(50, 50, 50), (100, 90, 100)])
def test_batch_with_deferred_blocks(ctx, ora, synth, pkg, q):
    """Frames that mix smooth content with noise: the noisy blocks have more symbols than the fast path takes, are queued
    and coded by heavy_blocks_kernel, and their chunks are woven back into the tiles by place_tiles_kernel."""
    torch = pytest.importorskip("torch")
    w, h, n = 512, 256, 3
    rng = np.random.default_rng(11)
    host = frames(synth, w, h, n, first=7).copy()
    for i in range(n):
        Y = host[i, : w * h].reshape(h, w)
        Y[:, 64 * i: 64 * i + 96] = rng.integers(0, 256, (h, 96), dtype=np.uint8)    # a noisy column band, moves per frame
        Y[100:108, :] = rng.integers(0, 256, (8, w), dtype=np.uint8)                 # and one noisy row of blocks
        U = host[i, w * h: w * h * 5 // 4].reshape(h // 2, w // 2)
        U[16:40, 8:200] = rng.integers(0, 256, (24, 192), dtype=np.uint8)
    d_in = torch.from_numpy(host).cuda()
    cap = pkg.capi.compress_bound(w, h) * n
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):  # twice: the queue and its counters are reused between launches
        ctx.compress_batch_dev(d_in, w, h, q, n, d_out, cap, d_off)
        ctx.batch_status()
        off, out = d_off.cpu().numpy(), d_out.cpu().numpy()
        for i in range(n):
            assert np.array_equal(out[off[i]: off[i + 1]], ora.compress(host[i], w, h, q)), f"frame {i}"
    d_back = torch.zeros_like(d_in)
    ctx.decompress_batch_dev(d_out, d_off, w, h, q, n, d_back)
    ctx.batch_status()
    for i in range(n):
        assert np.array_equal(d_back[i].cpu().numpy(), ora.decompress(out[off[i]: off[i + 1]], w, h, q))


def test_encoder_builds_produce_the_same_bytes(pkg, ora, synth):
    """The coding kernel exists in two builds (queue blocks above 8 symbols / code up to 15 in place, kernels.cu); which one a
    launch uses is a performance choice made from the previous launch's statistics and must never show in the bytes.  Mixed
    content (smooth frames with noisy stripes: blocks of 1..64 symbols) under both forced modes and under the automatic
    mode while it switches (three launches of detailed content, then three of smooth content)."""
    torch = pytest.importorskip("torch")
    w, h, n, q = 512, 256, 3, (75, 75, 75)
    rng = np.random.default_rng(12)
    mixed = frames(synth, w, h, n, first=2).copy()
    for i in range(n):
        Y = mixed[i, : w * h].reshape(h, w)
        Y[:, 32 * i: 32 * i + 200] = rng.integers(0, 256, (h, 200), dtype=np.uint8)
        Y[64:128, :] = (Y[64:128, :].astype(np.int32) + rng.integers(-20, 21, (64, w))).clip(0, 255).astype(np.uint8)
    smooth = synth.iyuv_frames_numpy(w, h, n, first=9, noise=False)
    cap = pkg.capi.compress_bound(w, h) * n
    c = pkg.Context(0)
    try:
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")

        def run(host):
            d_in = torch.from_numpy(host).cuda()
            torch.cuda.synchronize()
            c.compress_batch_dev(d_in, w, h, q, n, d_out, cap, d_off)
            c.batch_status()
            off = d_off.cpu().numpy()
            out = d_out.cpu().numpy()
            return [out[off[i]: off[i + 1]].copy() for i in range(n)]

        want_mixed = [ora.compress(mixed[i], w, h, q) for i in range(n)]
        want_smooth = [ora.compress(smooth[i], w, h, q) for i in range(n)]
        for mode in (1, 2):
            c.set_encoder_mode(mode)
            for got, want in zip(run(mixed), want_mixed):
                assert np.array_equal(got, want), f"mode {mode}"
            for got, want in zip(run(smooth), want_smooth):
                assert np.array_equal(got, want), f"mode {mode}"
        c.set_encoder_mode(0)
        for host, want in [(mixed, want_mixed)] * 3 + [(smooth, want_smooth)] * 3 + [(mixed, want_mixed)]:
            for got, wnt in zip(run(host), want):
                assert np.array_equal(got, wnt)
    finally:
        c.close()


def test_flat_frames_all_zero_blocks(ctx, ora):
    w, h = 64, 48 * 2  # 6144 px
    for level in (0, 128, 255):
        f = np.full(w * h * 3 // 2, level, np.uint8)
        got = ctx.compress(f, w, h, (50, 50, 50))
        assert np.array_equal(got, ora.compress(f, w, h, (50, 50, 50)))
        assert np.array_equal(ctx.decompress(got, w, h, (50, 50, 50)), ora.decompress(got, w, h, (50, 50, 50)))


@pytest.mark.parametrize("w,h", SMALL)
@pytest.mark.parametrize("q", QS[:3])
def test_decompress_matches_oracle(ctx, ora, synth, w, h, q):
    f = frames(synth, w, h, first=7)[0]
    payload = ora.compress(f, w, h, q)
    got = ctx.decompress(payload, w, h, q)
    assert np.array_equal(got, ora.decompress(payload, w, h, q))


def test_golden_files(ctx, golden_dir):
    """The reference's own golden vectors: bmp -> myyuv -> DCT-50 / DCT-90, decodes."""
    import oracle as O

    bmp = O.read_bmp32(golden_dir / "chef-with-trumpet.bmp")
    raw = O.read_myyuv(golden_dir / "chef-with-trumpet.myyuv")
    assert np.array_equal(ctx.xrgb_to_iyuv(bmp["data"], bmp["w"], bmp["h"], bmp["bottom_up"]), raw["data"])
    for q, dec_sha in ((50, "a95127da47152"), (90, "749ef0edb7ddd")):
        g = O.read_myyuv(golden_dir / f"chef-with-trumpet-DCT-{q}.myyuv")
        assert np.array_equal(ctx.compress(raw["data"], raw["w"], raw["h"], [q] * 3), g["data"])
        dec = ctx.decompress(g["data"], g["w"], g["h"], g["params"])
        hdr = O.YUV_HDR.pack(b"YU", 0x56555949, dec.size, 0, 0, 0, g["w"], g["h"], 64, bytes(32))
        assert hashlib.sha256(hdr + dec.tobytes()).hexdigest().startswith(dec_sha)


def test_golden_big_decode_and_recompress(ctx, ora, golden_dir):
    import oracle as O

    big = O.read_myyuv(golden_dir / "chef-with-trumpet-big-DCT-50.myyuv")
    dec = ctx.decompress(big["data"], big["w"], big["h"], big["params"])
    hdr = O.YUV_HDR.pack(b"YU", 0x56555949, dec.size, 0, 0, 0, big["w"], big["h"], 64, bytes(32))
    assert hashlib.sha256(hdr + dec.tobytes()).hexdigest().startswith("5e77691911882")
    # BASELINE config 2: 4032x3008 at q=90, and re-encoding at q=50 reproduces an oracle stream
    for q in (90, 50):
        got = ctx.compress(dec, big["w"], big["h"], [q] * 3)
        assert np.array_equal(got, ora.compress(dec, big["w"], big["h"], [q] * 3))


def test_cross_decode_with_reference(ctx, ref, synth):
    w, h, q = 320, 240 - 240 % 16, (60, 40, 80)
    f = frames(synth, w, h, first=11)[0]
    ours = ctx.compress(f, w, h, q)
    theirs = ref.compress(f, w, h, q)
    assert np.array_equal(ours, theirs)
    assert np.array_equal(ref.decompress(ours, w, h, q), ctx.decompress(theirs, w, h, q))


def test_batch_device_api(ctx, ora, synth, pkg):
    torch = pytest.importorskip("torch")
    w, h, n, q = 256, 144 - 144 % 16, 5, (50, 50, 50)
    host = frames(synth, w, h, n)
    d_in = torch.from_numpy(host).cuda()
    cap = pkg.capi.compress_bound(w, h) * n
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.compress_batch_dev(d_in, w, h, q, n, d_out, cap, d_off)
    ctx.batch_status()
    off = d_off.cpu().numpy()
    out = d_out.cpu().numpy()
    assert off[0] == 0
    for i in range(n):
        want = ora.compress(host[i], w, h, q)
        assert np.array_equal(out[off[i]: off[i + 1]], want), f"frame {i}"
    d_back = torch.zeros_like(d_in)
    ctx.decompress_batch_dev(d_out, d_off, w, h, q, n, d_back)
    ctx.batch_status()
    back = d_back.cpu().numpy()
    for i in range(n):
        assert np.array_equal(back[i], ora.decompress(out[off[i]: off[i + 1]], w, h, q))
    # colour conversion batch
    bg = synth.bgrx_frames_numpy(w, h, 3)
    d_bg = torch.from_numpy(bg).cuda()
    d_yuv = torch.empty((3, w * h * 3 // 2), dtype=torch.uint8, device="cuda")
    ctx.xrgb_to_iyuv_batch_dev(d_bg, w, h, True, 3, d_yuv)
    ctx.batch_status()
    got = d_yuv.cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], ora.bgrx_to_iyuv(bg[i], w, h, True))


@pytest.mark.parametrize("content", ["gradient", "noise"])
@pytest.mark.parametrize("w,h,n,q", [(48, 80, 3, (50, 50, 50)), (1296, 720, 2, (90, 90, 90)), (256, 128, 4, (100, 100, 100))])
def test_no_write_outside_the_callers_buffers(ctx, ora, synth, pkg, content, w, h, n, q):
    """compute-sanitizer is not available on the pool, so the bounds of the tile / scan / place / finalize kernels and of the
    decoder's stores are checked with guard bands: every device buffer the call writes sits between two 4 KB bands of a known
    byte, the payload buffer is exactly as large as the payload, and the bands must come back untouched."""
    torch = pytest.importorskip("torch")
    G, fill = 4096, 0xA5
    rng = np.random.default_rng(w * h + n)
    host = frames(synth, w, h, n) if content == "gradient" else rng.integers(0, 256, (n, w * h * 3 // 2), dtype=np.uint8)
    want = [ora.compress(host[i], w, h, q) for i in range(n)]
    total = sum(len(x) for x in want)

    def banded(nbytes, dtype=torch.uint8):
        raw = torch.full((G + nbytes + G,), fill, dtype=torch.uint8, device="cuda")
        return raw, raw[G: G + nbytes].view(dtype)

    def bands_intact(raw, nbytes):
        r = raw.cpu().numpy()
        return bool((r[:G] == fill).all() and (r[G + nbytes:] == fill).all())

    d_in = torch.from_numpy(host).cuda()
    raw_out, d_out = banded(total)
    raw_off, d_off = banded(8 * (n + 1), torch.int64)
    torch.cuda.synchronize()
    for mode in (1, 2):  # both builds of the coding kernel
        ctx.set_encoder_mode(mode)
        ctx.compress_batch_dev(d_in, w, h, q, n, d_out, total, d_off)
        ctx.batch_status()
        assert bands_intact(raw_out, total) and bands_intact(raw_off, 8 * (n + 1)), f"encoder mode {mode}"
        off = d_off.cpu().numpy()
        assert off[n] == total
        out = d_out.cpu().numpy()
        for i in range(n):
            assert np.array_equal(out[off[i]: off[i + 1]], want[i])
    ctx.set_encoder_mode(0)
    raw_back, d_back = banded(host.size)
    ctx.decompress_batch_dev(d_out, d_off, w, h, q, n, d_back.view(n, -1))
    ctx.batch_status()
    assert bands_intact(raw_back, host.size)
    back = d_back.cpu().numpy().reshape(n, -1)
    for i in range(n):
        assert np.array_equal(back[i], ora.decompress(want[i], w, h, q))
    # one byte short: the call must report it and still leave the bands alone
    raw_short, d_short = banded(total - 1)
    ctx.compress_batch_dev(d_in, w, h, q, n, d_short, total - 1, d_off)
    with pytest.raises(pkg.MyyuvError):
        ctx.batch_status()
    assert bands_intact(raw_short, total - 1)


@pytest.mark.parametrize("chunk,keep_iyuv", [(0, False), (1, True), (3, False), (4, True), (7, True)])
def test_full_pipeline_xrgb_to_payload(ctx, ora, synth, pkg, chunk, keep_iyuv):
    """BASELINE configs[2]: XRGB -> IYUV -> DCT-50 in one call, chunks chained on the device; every frame's payload and
    (optionally) IYUV image must equal the oracle's two-step result."""
    torch = pytest.importorskip("torch")
    w, h, n, q = 320, 176, 7, (50, 50, 50)
    bg = synth.bgrx_frames_numpy(w, h, n, first=2)
    d_bg = torch.from_numpy(bg).cuda()
    cap = pkg.capi.compress_bound(w, h) * n
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.full((n + 1,), -1, dtype=torch.int64, device="cuda")
    d_yuv = torch.zeros((n, w * h * 3 // 2), dtype=torch.uint8, device="cuda") if keep_iyuv else None
    torch.cuda.synchronize()
    ctx.xrgb_compress_batch_dev(d_bg, w, h, True, q, n, d_out, cap, d_off, d_yuv, chunk)
    ctx.batch_status()
    off, out = d_off.cpu().numpy(), d_out.cpu().numpy()
    assert off[0] == 0 and np.all(np.diff(off) > 0)
    for i in range(n):
        iyuv = ora.bgrx_to_iyuv(bg[i], w, h, True)
        if keep_iyuv:
            assert np.array_equal(d_yuv[i].cpu().numpy(), iyuv), f"frame {i} iyuv"
        assert np.array_equal(out[off[i]: off[i + 1]], ora.compress(iyuv, w, h, q)), f"frame {i}"
    # a capacity that cuts the batch short is reported, nothing is written past it
    small = int(off[4]) + 10
    d_out2 = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    ctx.xrgb_compress_batch_dev(d_bg, w, h, True, q, n, d_out2, small, d_off, None, chunk)
    with pytest.raises(pkg.MyyuvError) as e:
        ctx.batch_status()
    assert e.value.code == pkg.capi.ERR_CAPACITY
    assert not d_out2[small:].any()


def test_batch_host_api_multi_chunk(ctx, ora, synth, pkg):
    # 1920x1088 frames are 3.1 MB: 30 frames span three 32 MB pipeline chunks
    w, h, n, q = 1920, 1088, 30, (50, 50, 50)
    host = frames(synth, w, h, n)
    cap = 40 << 20
    out = np.empty(cap, np.uint8)
    off = np.zeros(n + 1, np.uint64)
    ctx.compress_batch_host(host, w, h, q, n, out, off)
    for i in (0, 1, 20, 21, 29):
        assert np.array_equal(out[int(off[i]): int(off[i + 1])], ora.compress(host[i], w, h, q)), f"frame {i}"
    back = np.empty_like(host)
    ctx.decompress_batch_host(out, off, w, h, q, n, back)
    for i in (0, 19, 20, 21, 29):
        assert np.array_equal(back[i], ora.decompress(out[int(off[i]): int(off[i + 1])], w, h, q))


def test_batch_host_api_pinned_buffers(ctx, ora, synth, pkg):
    """Pinned (mapped) caller buffers take the SM-copy path for payloads and offsets; payloads start at arbitrary byte
    positions of the caller's buffer, so the copy kernel's unaligned cases are all met.  Same bytes as pageable buffers."""
    w, h, n, q = 1920, 1088, 24, (50, 50, 50)
    host = frames(synth, w, h, n)
    fb = w * h * 3 // 2
    pin_in, pin_out, pin_back = pkg.capi.PinnedBuffer(n * fb), pkg.capi.PinnedBuffer((30 << 20) + 3), pkg.capi.PinnedBuffer(n * fb)
    pin_in.array[:] = host.reshape(-1)
    out_p, off_p = pin_out.array[3:], np.zeros(n + 1, np.uint64)      # odd start address
    ctx.compress_batch_host(pin_in.array, w, h, q, n, out_p, off_p)
    out, off = np.empty(30 << 20, np.uint8), np.zeros(n + 1, np.uint64)
    ctx.compress_batch_host(host, w, h, q, n, out, off)
    assert np.array_equal(off, off_p)
    total = int(off[n])
    assert np.array_equal(out[:total], out_p[:total])
    for i in (0, 9, 10, 23):
        assert np.array_equal(out_p[int(off[i]): int(off[i + 1])], ora.compress(host[i], w, h, q)), f"frame {i}"
    ctx.decompress_batch_host(out_p, off_p, w, h, q, n, pin_back.array)
    back = np.empty_like(host)
    ctx.decompress_batch_host(out, off, w, h, q, n, back)
    assert np.array_equal(back.reshape(-1), pin_back.array)
    assert np.array_equal(back[11], ora.decompress(out[int(off[11]): int(off[12])], w, h, q))
    # a payload buffer that is too small is still reported, and the context stays usable
    with pytest.raises(pkg.MyyuvError):
        ctx.compress_batch_host(pin_in.array, w, h, q, n, pin_out.array[: total // 2], off_p)
    ctx.compress_batch_host(pin_in.array, w, h, q, n, out_p, off_p)
    assert np.array_equal(out[:total], out_p[:total])


@pytest.mark.parametrize("env", [{"MYYUVB_SMALL_COPY": "0", "MYYUVB_CHUNK_MB": "4"}, {"MYYUVB_D2H_STREAM": "0", "MYYUVB_CHUNK_MB": "4"},
                                 {"MYYUVB_CHUNK_MB": "4", "MYYUVB_STAGING": "direct"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_batch_host_api_env_switches(env, tmp_path):
    """The switches of INTEGRATION.md select other copy paths of the *_batch_host calls (read once per process, hence a
    subprocess); every one must give the bytes of the default path, here checked against the oracle."""
    import os, pathlib, subprocess, sys
    root = pathlib.Path(__file__).resolve().parent.parent
    script = tmp_path / "switches.py"
    script.write_text('''
import importlib, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import oracle
pkg = importlib.import_module("yuv-manipulations-2_b200"); synth = importlib.import_module("yuv-manipulations-2_b200.synth")
w, h, n, q = 1280, 720 - 720 % 16, 9, (60, 40, 70)
fb = w * h * 3 // 2
host = synth.iyuv_frames_numpy(w, h, n, 11)
ctx = pkg.Context(0)
ora = oracle.Oracle()
for pinned in (True, False):
    if pinned:
        a, b, c = pkg.capi.PinnedBuffer(n * fb), pkg.capi.PinnedBuffer((8 << 20) + 1), pkg.capi.PinnedBuffer(n * fb)
        a.array[:] = host.reshape(-1)
        src, out, back = a.array, b.array[1:], c.array
    else:
        src, out, back = host.reshape(-1).copy(), np.empty(8 << 20, np.uint8), np.empty(n * fb, np.uint8)
    off = np.zeros(n + 1, np.uint64)
    ctx.compress_batch_host(src, w, h, q, n, out, off)
    for i in (0, 4, 8):
        assert np.array_equal(out[int(off[i]): int(off[i + 1])], ora.compress(host[i], w, h, q)), (pinned, i)
    ctx.decompress_batch_host(out, off, w, h, q, n, back)
    assert np.array_equal(back[8 * fb:], ora.decompress(out[int(off[8]): int(off[9])], w, h, q)), pinned
    assert np.array_equal(back[:fb], ora.decompress(out[int(off[0]): int(off[1])], w, h, q)), pinned
print("switches ok")
''')
    r = subprocess.run([sys.executable, str(script), str(root)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "switches ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def _golden():
    import json
    import pathlib

    return json.loads((pathlib.Path(__file__).resolve().parent / "golden" / "golden.json").read_text())


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_headline_4k_frame_matches_reference_hashes(ctx, synth):
    """The benchmark's own frame 0 (3840x2160 noise-grad, q 50): payload and decoded image against the hashes the
    unmodified reference produced for it (tests/golden/make_golden.py)."""
    case = next(c for c in _golden()["synthetic"] if c["w"] == 3840)
    w, h, q = case["w"], case["h"], tuple(case["q"])
    f = frames(synth, w, h, 1, case["first"])[0]
    assert _sha(f) == case["input_sha256"], "synthetic generator changed"
    p = ctx.compress(f, w, h, q)
    assert p.size == case["payload_size"]
    assert _sha(p) == case["payload_sha256"]
    assert _sha(ctx.decompress(p, w, h, q)) == case["decoded_sha256"]


def test_headline_4k_batch_of_64_matches_oracle(ctx, ora, synth, pkg):
    """bench.py's step -- 64 frames of 3840x2160 at q 50 through compress_batch_dev / decompress_batch_dev -- with frames
    0, 31 and 63 compared byte for byte with the oracle (frame 0 also with the reference's hash)."""
    torch = pytest.importorskip("torch")
    w, h, n, q = 3840, 2160, 64, (50, 50, 50)
    d_in = synth.iyuv_frames_torch(w, h, n, torch.device("cuda", 0))
    cap = n * 6 * 1024 * 1024
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    d_back = torch.empty_like(d_in)
    torch.cuda.synchronize()
    ctx.compress_batch_dev(d_in, w, h, q, n, d_out, cap, d_off)
    ctx.decompress_batch_dev(d_out, d_off, w, h, q, n, d_back)
    ctx.batch_status()
    off = d_off.cpu().numpy()
    assert off[0] == 0 and (np.diff(off) > 0).all()
    case = next(c for c in _golden()["synthetic"] if c["w"] == 3840)
    for i in (0, 31, 63):
        f = d_in[i].cpu().numpy()
        assert np.array_equal(f, frames(synth, w, h, 1, i)[0]), "torch and numpy generators differ"
        got = d_out[int(off[i]): int(off[i + 1])].cpu().numpy()
        want = ora.compress(f, w, h, q)
        assert got.size == want.size and np.array_equal(got, want), f"payload of frame {i}"
        assert np.array_equal(d_back[i].cpu().numpy(), ora.decompress(want, w, h, q)), f"decoded frame {i}"
        if i == 0:
            assert _sha(got) == case["payload_sha256"] and _sha(d_back[0].cpu().numpy()) == case["decoded_sha256"]
    del d_in, d_out, d_back
    torch.cuda.empty_cache()


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_natural_content_4k_matches_reference_and_oracle(ctx, ora, synth, golden_dir, idx):
    """Natural content at the benchmark's frame size (the reference's sample image tiled to 4K)
    at q 50, 90 and 10: against the reference's hashes and, byte for byte, the oracle."""
    import oracle

    case = _golden()["tiled_real"][idx]
    g = oracle.read_myyuv(golden_dir / "chef-with-trumpet.myyuv")
    w, h, q = case["w"], case["h"], tuple(case["q"])
    f = synth.tiled_real_iyuv(g["data"], g["w"], g["h"], w, h, 1, case["first"])[0]
    assert _sha(f) == case["input_sha256"]
    p = ctx.compress(f, w, h, q)
    assert p.size == case["payload_size"] and _sha(p) == case["payload_sha256"]
    assert np.array_equal(p, ora.compress(f, w, h, q))
    d = ctx.decompress(p, w, h, q)
    assert _sha(d) == case["decoded_sha256"]
    assert np.array_equal(d, ora.decompress(p, w, h, q))


def test_round_trip_properties_4k(ctx, synth):
    """Size-independent properties at the benchmark's frame size: decode(encode(x)) is idempotent under a
    second encode/decode cycle's stream sizes, planes stay within the quantisation error, chunk sizes sum up."""
    w, h, q = 3840, 2160, (50, 50, 50)
    f = frames(synth, w, h)[0]
    p1 = ctx.compress(f, w, h, q)
    d1 = ctx.decompress(p1, w, h, q)
    psz = p1[:12].view(np.uint32)
    assert 12 + int(psz.sum()) == p1.size
    pos = 12
    for p, n in enumerate((w * h // 64, w * h // 256, w * h // 256)):
        nchunks, content = p1[pos: pos + 8].view(np.uint32)
        assert nchunks == n and psz[p] == 8 + n + content
        assert int(p1[pos + 8: pos + 8 + n].sum(dtype=np.int64)) == content
        pos += int(psz[p])
    err = np.abs(d1.astype(np.int16) - f.astype(np.int16))
    assert err.max() < 64 and err.mean() < 6
    p2 = ctx.compress(d1, w, h, q)
    d2 = ctx.decompress(p2, w, h, q)
    assert np.abs(d2.astype(np.int16) - d1.astype(np.int16)).mean() < 1.0


def test_error_behaviour(ctx, pkg, ora, synth):
    M = pkg.MyyuvError
    f = frames(synth, 32, 32)[0]
    with pytest.raises(M, match="Level of quality must be between 1 and 100"):
        ctx.compress(f, 32, 32, (0, 50, 50))
    with pytest.raises(M, match="Level of quality must be between 1 and 100"):
        ctx.compress(f, 32, 32, (50, 101, 50))
    with pytest.raises(M, match="width % 8 must be 0"):
        ctx.compress(np.zeros(24 * 32 * 3 // 2, np.uint8), 24, 32, (50, 50, 50))  # chroma width 12
    with pytest.raises(M, match="height % 8 must be 0"):
        ctx.compress(np.zeros(32 * 24 * 3 // 2, np.uint8), 32, 24, (50, 50, 50))
    good = ora.compress(f, 32, 32, (50, 50, 50))
    with pytest.raises(M, match="DCTYUV load bad size"):
        ctx.decompress(good[:12], 32, 32, (50, 50, 50))
    with pytest.raises(M, match="DCTYUV load bad size"):
        ctx.decompress(good[:-1], 32, 32, (50, 50, 50))
    bad = good.copy()
    bad[12:16] = 0  # n_chunks = 0
    with pytest.raises(M, match="DCTYUVPlane load"):
        ctx.decompress(bad, 32, 32, (50, 50, 50))
    with pytest.raises(M, match="too small"):
        ctx.compress(f, 32, 32, (50, 50, 50), capacity=good.size - 1)
    # context still usable afterwards
    assert np.array_equal(ctx.compress(f, 32, 32, (50, 50, 50)), good)


def test_corrupt_code_stream_is_reported(ctx, ora, synth):
    w, h, q = 64, 64, (90, 90, 90)
    f = frames(synth, w, h)[0]
    good = ora.compress(f, w, h, q)
    n = w * h // 64
    first_chunk = 12 + 8 + n
    bad = good.copy()
    bad[first_chunk] = 0xFF  # code_bits low byte: far more bits than the chunk holds
    bad[first_chunk + 1] = 0x01
    with pytest.raises(Exception, match="Huffman bad code"):
        ctx.decompress(bad, w, h, q)
    with pytest.raises(Exception):
        ora.decompress(bad, w, h, q)


def test_damaged_chunks_gpu_decoder_agrees_with_its_host_build(ctx, ora, synth, pkg):
    """The decoder kernel's table parser and stream loop are written in PTX for the kernel's shared-memory layout; the same
    decoder (block_codec.cuh) compiled for the host is what tests/test_hostemu.py checks against the oracle.  Random byte flips
    in the chunks of the luma plane: the kernel must flag an error exactly when the host build does, and where both accept the
    chunks, the pixels must be those of the coefficients the host build decoded (re-encoded cleanly and decoded by the oracle)."""
    import ctypes as C
    import pathlib
    import subprocess

    here = pathlib.Path(__file__).parent / "hostemu"
    subprocess.run(["make", "-s", "-C", str(here)], check=True)
    emu = C.CDLL(str(here / "libhostemu.so"))
    u8p, i16p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16)
    emu.hostemu_decode_blocks2.argtypes = [u8p, u8p, C.c_uint32, C.c_int, i16p, C.POINTER(C.c_uint32)]
    w, h, q = 64, 64, (90, 90, 90)
    rng = np.random.default_rng(20261019)
    f = rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8)  # noise: tables of every kind, long streams
    f[: w * h // 2] = frames(synth, w, h)[0][: w * h // 2]     # and smooth content in the upper half
    good = np.asarray(ora.compress(f, w, h, q), np.uint8)
    psz = good[:12].view(np.uint32)
    n = int(good[12:16].view(np.uint32)[0])
    content_size = int(good[16:20].view(np.uint32)[0])
    sizes_at, content_at = 20, 20 + n
    assert n == (w // 8) * (h // 8) and psz[0] == 8 + n + content_size
    sizes = np.ascontiguousarray(good[sizes_at: sizes_at + n])
    rest = good[12 + int(psz[0]):]  # the two chroma planes, untouched
    errors = accepted = 0
    for trial in range(120):
        bad = good.copy()
        for _ in range(int(rng.integers(1, 4))):
            at = content_at + int(rng.integers(0, content_size))
            bad[at] ^= np.uint8(rng.integers(1, 256))
        chunks = np.ascontiguousarray(bad[content_at: content_at + content_size])
        coefs = np.zeros(n * 64, np.int16)
        err = emu.hostemu_decode_blocks2(chunks.ctypes.data_as(u8p), sizes.ctypes.data_as(u8p), n, 1, coefs.ctypes.data_as(i16p), None)
        try:
            got = ctx.decompress(bad, w, h, q)
            gpu_err = False
        except pkg.MyyuvError:
            gpu_err = True
        assert gpu_err == (err != 0), f"trial {trial}: host build says {err}, kernel says {gpu_err}"
        if err:
            errors += 1
            continue
        accepted += 1
        ch, sz = ora.huff_encode_blocks(coefs.reshape(n, 64))
        ch, sz = np.asarray(ch, np.uint8).reshape(-1), np.asarray(sz, np.uint8).reshape(-1)
        head = np.array([8 + n + ch.size, psz[1], psz[2]], np.uint32).view(np.uint8)
        plane = np.concatenate([np.array([n, ch.size], np.uint32).view(np.uint8), sz, ch])
        clean = np.concatenate([head, plane, rest])
        assert np.array_equal(got, ora.decompress(clean, w, h, q)), f"trial {trial}"
    assert errors >= 10 and accepted >= 10, (errors, accepted)


def test_class_api_round_trip(pkg, ora, synth, tmp_path):
    """The reference-style class API (YUV(bmp, IYUV) -> compress -> dump -> load -> decompress)."""
    YUV, BMP = pkg.YUV, pkg.BMP
    w, h = 64, 48 - 48 % 16
    bgrx = synth.bgrx_frames_numpy(w, h, 1)[0]
    bmp = BMP()
    bmp.header.width, bmp.header.height, bmp.header.bit_count = w, h, 32
    bmp.header.header_size, bmp.header.compression, bmp.header.planes = 124, 3, 1
    bmp.header.data_pos = 138
    bmp.header.file_size = 138 + bgrx.size
    bmp.data = bgrx.reshape(-1).copy()
    bmp.dump(str(tmp_path / "a.bmp"))
    bmp2 = BMP(str(tmp_path / "a.bmp"))
    yuv = YUV(bmp2, YUV.FourccFormats.IYUV)
    assert yuv.isValid() and not yuv.isCompressed()
    assert np.array_equal(yuv.data, ora.bgrx_to_iyuv(bgrx, w, h, True))
    c = yuv.compress(YUV.Compressions.DCT, [50, 60, 70])
    assert c.isValid() and c.isCompressed() and c.header.data_pos == 67 and c.header.compression_params_size == 3
    assert np.array_equal(c.data, ora.compress(yuv.data, w, h, (50, 60, 70)))
    c.dump(str(tmp_path / "c.myyuv"))
    c2 = YUV(str(tmp_path / "c.myyuv"))
    d = c2.decompress()
    assert d.isValid() and d.header.data_size == w * h * 3 // 2 and d.header.data_pos == 64
    assert np.array_equal(d.data, ora.decompress(c.data, w, h, (50, 60, 70)))
    with pytest.raises(RuntimeError, match="Error already compressed"):
        c.compress(YUV.Compressions.DCT, [50, 50, 50])
    with pytest.raises(RuntimeError, match="3 parameters required"):
        yuv.compress(YUV.Compressions.DCT, [50])
