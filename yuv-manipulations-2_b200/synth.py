"""Deterministic synthetic frames (SURVEY.md section 8(d) "noise-grad"): integer-only formulas, so the
numpy (CPU, tests / cpu_baseline) and torch (GPU, bench) versions produce identical bytes.

  v(x, y, f, c) = clamp( tri(3x + 5f + 37c, 512)/2 + tri(2y + 91c, 384)/2 + ((h32(x, y, f, c) & 15) - 8), 0, 255 )
  tri(t, P)     = |t mod P - P/2| * 510 / P            (integer division)
  h32           = murmur3 finaliser of  x + 7919*y + 104729*f + 1000003*c + seed   (mod 2^32)
"""
from __future__ import annotations

import numpy as np

SEED = 20261018
_M = 0xFFFFFFFF


def _h32(v, xp):
    v = v & _M
    v = v ^ (v >> 16)
    v = (v * 0x85EBCA6B) & _M
    v = v ^ (v >> 13)
    v = (v * 0xC2B2AE35) & _M
    v = v ^ (v >> 16)
    return v


def _plane(xp, w, h, f, c, seed, noise=True, **kw):
    # int64 arithmetic everywhere (torch has no uint32 ops)
    x = xp.arange(w, dtype=xp.int64, **kw).reshape(1, w)
    y = xp.arange(h, dtype=xp.int64, **kw).reshape(h, 1)

    def tri(t, P):
        return abs((t % P) - P // 2) * 510 // P

    v = tri(3 * x + 5 * f + 37 * c, 512) // 2 + tri(2 * y + 91 * c, 384) // 2
    if noise:
        hsh = _h32(x + 7919 * y + (104729 * f + 1000003 * c + seed), xp)
        v = v + (hsh & 15) - 8
    else:
        v = v + 0 * y
    return v.clip(0, 255)


def iyuv_frames_numpy(w: int, h: int, n_frames: int, first: int = 0, seed: int = SEED, noise: bool = True) -> np.ndarray:
    """[n_frames, w*h*3/2] uint8: Y (c=0) at full resolution, U (c=1) and V (c=2) at half resolution."""
    out = np.empty((n_frames, w * h * 3 // 2), np.uint8)
    for i in range(n_frames):
        f = first + i
        out[i, : w * h] = _plane(np, w, h, f, 0, seed, noise).astype(np.uint8).reshape(-1)
        out[i, w * h: w * h * 5 // 4] = _plane(np, w // 2, h // 2, f, 1, seed, noise).astype(np.uint8).reshape(-1)
        out[i, w * h * 5 // 4:] = _plane(np, w // 2, h // 2, f, 2, seed, noise).astype(np.uint8).reshape(-1)
    return out


def iyuv_frames_torch(w: int, h: int, n_frames: int, device, first: int = 0, seed: int = SEED, noise: bool = True):
    import torch

    out = torch.empty((n_frames, w * h * 3 // 2), dtype=torch.uint8, device=device)
    for i in range(n_frames):
        f = first + i
        out[i, : w * h] = _plane(torch, w, h, f, 0, seed, noise, device=device).to(torch.uint8).reshape(-1)
        out[i, w * h: w * h * 5 // 4] = _plane(torch, w // 2, h // 2, f, 1, seed, noise, device=device).to(torch.uint8).reshape(-1)
        out[i, w * h * 5 // 4:] = _plane(torch, w // 2, h // 2, f, 2, seed, noise, device=device).to(torch.uint8).reshape(-1)
    return out


def bgrx_frames_numpy(w: int, h: int, n_frames: int, first: int = 0, seed: int = SEED) -> np.ndarray:
    """[n_frames, h, w, 4] uint8 B,G,R,X (X = 0), rows in file order."""
    out = np.zeros((n_frames, h, w, 4), np.uint8)
    for i in range(n_frames):
        for c in range(3):
            out[i, :, :, c] = _plane(np, w, h, first + i, c, seed).astype(np.uint8)
    return out


def bgrx_frames_torch(w: int, h: int, n_frames: int, device, first: int = 0, seed: int = SEED):
    import torch

    out = torch.zeros((n_frames, h, w, 4), dtype=torch.uint8, device=device)
    for i in range(n_frames):
        for c in range(3):
            out[i, :, :, c] = _plane(torch, w, h, first + i, c, seed, device=device).to(torch.uint8)
    return out


def edge_case_iyuv(w: int, h: int) -> np.ndarray:
    """One frame mixing the adversarial cases of SURVEY 8(d)(iii): flat 128 (all-zero blocks), 8x8 checkerboards
    (many distinct symbols), saturated stripes, a noise band; chroma gets ramps and extremes."""
    rng = np.random.default_rng(SEED)
    Y = np.full((h, w), 128, np.uint8)
    q = h // 4
    yy, xx = np.mgrid[0:q, 0:w]
    Y[q: 2 * q] = (((yy + xx) & 1) * 255).astype(np.uint8)
    Y[2 * q: 3 * q] = ((xx // 3) % 2 * 255).astype(np.uint8)
    Y[3 * q: 3 * q + q] = rng.integers(0, 256, (h - 3 * q, w), dtype=np.uint8)[:q]
    U = np.tile(np.linspace(0, 255, w // 2).astype(np.uint8), (h // 2, 1))
    V = np.where((np.mgrid[0:h // 2, 0:w // 2][0] // 4) % 2 == 0, 0, 255).astype(np.uint8)
    return np.concatenate([Y.reshape(-1), U.reshape(-1), V.reshape(-1)])


def tiled_real_iyuv(base: np.ndarray, bw: int, bh: int, w: int, h: int, n_frames: int, first: int = 0) -> np.ndarray:
    """SURVEY 8(d)(i) "tiled-real": frame f = a natural IYUV image (bw x bh, e.g. the reference's chef-with-trumpet.myyuv)
    tiled to w x h with its origin shifted by (16 f mod bw, 16 f mod bh).  [n_frames, w*h*3/2] uint8."""
    Y = base[: bw * bh].reshape(bh, bw)
    U = base[bw * bh: bw * bh * 5 // 4].reshape(bh // 2, bw // 2)
    V = base[bw * bh * 5 // 4:].reshape(bh // 2, bw // 2)
    out = np.empty((n_frames, w * h * 3 // 2), np.uint8)
    for i in range(n_frames):
        f = first + i
        sx, sy = (16 * f) % bw, (16 * f) % bh
        yy = (np.arange(h) + sy) % bh
        xx = (np.arange(w) + sx) % bw
        out[i, : w * h] = Y[np.ix_(yy, xx)].reshape(-1)
        yc = (np.arange(h // 2) + sy // 2) % (bh // 2)
        xc = (np.arange(w // 2) + sx // 2) % (bw // 2)
        out[i, w * h: w * h * 5 // 4] = U[np.ix_(yc, xc)].reshape(-1)
        out[i, w * h * 5 // 4:] = V[np.ix_(yc, xc)].reshape(-1)
    return out
