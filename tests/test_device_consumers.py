"""SURVEY 8(f) rows 1 (second half) and 4: what the reference's viewers do with a decoded frame on the host -- plane
accessors, YUV::getPixel, the fragment shader's YUV -> RGB -- on device memory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rgba_reference(iyuv, w, h):
    """frag_yuv.glsl:18-26 in double precision, chroma sampled bilinearly (GL_LINEAR, clamp to edge) at the luma pixel
    centres of a 1:1 display; 8-bit output = round(clamp(c, 0, 1) * 255)."""
    Y = iyuv[: w * h].reshape(h, w).astype(np.float64)
    U = iyuv[w * h: w * h * 5 // 4].reshape(h // 2, w // 2).astype(np.float64)
    V = iyuv[w * h * 5 // 4:].reshape(h // 2, w // 2).astype(np.float64)

    def up(P):
        ch, cw = P.shape
        yy = (np.arange(h) + 0.5) / 2 - 0.5
        xx = (np.arange(w) + 0.5) / 2 - 0.5
        y0 = np.floor(yy).astype(int)
        x0 = np.floor(xx).astype(int)
        fy, fx = (yy - y0)[:, None], (xx - x0)[None, :]
        ya, yb = np.clip(y0, 0, ch - 1), np.clip(y0 + 1, 0, ch - 1)
        xa, xb = np.clip(x0, 0, cw - 1), np.clip(x0 + 1, 0, cw - 1)
        top = P[np.ix_(ya, xa)] * (1 - fx) + P[np.ix_(ya, xb)] * fx
        bot = P[np.ix_(yb, xa)] * (1 - fx) + P[np.ix_(yb, xb)] * fx
        return top * (1 - fy) + bot * fy

    y, u, v = Y / 255.0, up(U) / 255.0 - 0.5, up(V) / 255.0 - 0.5
    rgb = np.stack([y + 1.403 * v, y - 0.714 * v - 0.344 * u, y + 1.773 * u], -1)
    out = np.full((h, w, 4), 255, np.uint8)
    out[..., :3] = np.rint(np.clip(rgb, 0, 1) * 255).astype(np.uint8)
    return out


@pytest.mark.parametrize("w,h", [(16, 16), (64, 48), (1920, 1088)])
def test_iyuv_to_rgba_matches_the_shader_formula(ctx, synth, w, h):
    torch = pytest.importorskip("torch")
    n = 2
    f = synth.iyuv_frames_numpy(w, h, n, first=4)
    rng = np.random.default_rng(7)
    f[1, w * h:] = rng.integers(0, 256, w * h // 2, dtype=np.uint8)  # saturated, noisy chroma: every clamp is exercised
    d_in = torch.from_numpy(f).cuda()
    d_out = torch.empty((n, h, w, 4), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for flip in (False, True):
        ctx.iyuv_to_rgba_batch_dev(d_in, w, h, n, d_out, flip)
        ctx.batch_status()
        got = d_out.cpu().numpy()
        for i in range(n):
            want = rgba_reference(f[i], w, h)
            g = got[i][::-1] if flip else got[i]
            diff = np.abs(g.astype(np.int16) - want.astype(np.int16))
            assert diff.max() <= 1, f"frame {i} flip {flip}: max |diff| {diff.max()}"   # tolerance: +-1 LSB (float32 vs float64 at rounding ties)
            assert (diff != 0).mean() < 0.01
            assert (g[..., 3] == 255).all()


def test_get_pixels_follow_the_reference_indexing(ctx, pkg, synth):
    """YUV::getPixel's IYUV entry (myyuv_yuv.cpp:162-180): Y at x + y * width, chroma at x / 2 + y * width / 4 in both planes
    (for odd y that is half a chroma row past row y / 2 -- reproduced, because viewers built on the reference read these bytes)."""
    torch = pytest.importorskip("torch")
    w, h = 64, 48
    f = synth.iyuv_frames_numpy(w, h, 1, first=1)[0]
    xs, ys = np.meshgrid(np.arange(w, dtype=np.uint32), np.arange(h, dtype=np.uint32))
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.uint32)
    d_f = torch.from_numpy(f).cuda()
    d_xy = torch.from_numpy(xy.view(np.int32)).cuda()
    d_out = torch.empty(xy.shape[0] * 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.get_pixels_dev(d_f, w, h, xy.shape[0], d_xy, d_out)
    ctx.batch_status()
    got = d_out.cpu().numpy().reshape(-1, 3)
    x, y = xy[:, 0].astype(np.int64), xy[:, 1].astype(np.int64)
    uv = x // 2 + y * w // 4
    vi = w * h * 5 // 4 + uv
    inside = vi < f.size  # on the last (odd) row the reference's index leaves its buffer for x >= w / 2: defined as 0 here
    assert (~inside).sum() == w // 2
    want = np.stack([f[x + y * w], f[w * h + uv], np.where(inside, f[np.minimum(vi, f.size - 1)], 0)], 1)
    assert np.array_equal(got, want)
    planes = pkg.capi.iyuv_planes(d_f.data_ptr(), w, h)
    assert planes == [(d_f.data_ptr(), w, h), (d_f.data_ptr() + w * h, w // 2, h // 2), (d_f.data_ptr() + w * h * 5 // 4, w // 2, h // 2)]
    bad = torch.tensor([[w, 0]], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.get_pixels_dev(d_f, w, h, 1, bad, d_out)
    with pytest.raises(pkg.MyyuvError, match="Image coordinates are out of bounds"):
        ctx.batch_status()


def test_decode_to_rgba_pipeline(ctx, ora, pkg, synth):
    """payloads -> RGBA in one call, chunked: the same bytes as decoding and converting separately, and the decoded frames
    the oracle gives."""
    torch = pytest.importorskip("torch")
    w, h, n, q = 256, 128, 5, (60, 60, 60)
    f = synth.iyuv_frames_numpy(w, h, n, first=2)
    pay = [ora.compress(f[i], w, h, q) for i in range(n)]
    off = np.concatenate([[0], np.cumsum([p.size for p in pay])]).astype(np.int64)
    d_pay = torch.from_numpy(np.concatenate(pay)).cuda()
    d_off = torch.from_numpy(off).cuda()
    d_rgba = torch.empty((n, h, w, 4), dtype=torch.uint8, device="cuda")
    d_iyuv = torch.empty((n, w * h * 3 // 2), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for chunk, keep in ((0, True), (2, False), (1, True)):
        d_rgba.zero_()
        torch.cuda.synchronize()
        ctx.decompress_to_rgba_batch_dev(d_pay, d_off, w, h, q, n, d_rgba, d_iyuv if keep else None, chunk)
        ctx.batch_status()
        got = d_rgba.cpu().numpy()
        for i in range(n):
            dec = ora.decompress(pay[i], w, h, q)
            if keep:
                assert np.array_equal(d_iyuv[i].cpu().numpy(), dec)
            diff = np.abs(got[i].astype(np.int16) - rgba_reference(dec, w, h).astype(np.int16))
            assert diff.max() <= 1
