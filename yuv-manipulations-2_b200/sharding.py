"""Multi-GPU sharding of the hot path; one process per GPU, torch.distributed for plumbing.

Two ways the path shards, both without any exchange while coding:

* a batch of frames: independent units, contiguous ranges per rank (``frame_range``), no collective at all;
* one very large image: every 8x8 block is coded independently (own Huffman table, byte aligned, DCT.cpp:297-313)
  and a plane's payload is the concatenation of its blocks in raster order, so a horizontal band of macroblock
  rows is a complete IYUV image of smaller height.  Rank r compresses its band with the ordinary kernel; the
  single exchange step is assembling the stream: all-gather of the band payload sizes, gather of the band
  payloads to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests), then a byte re-arrangement
  (``assemble_payload``): per plane, size arrays of all bands followed by contents of all bands.

Decompression shards the same way: rank 0 cuts the payload into per-band payloads (``split_payload``) using the
prefix sum of the chunk sizes, scatters them, every rank decodes its band.

The coder itself is passed in (``compress_fn`` / ``decompress_fn``): the product passes the CUDA context's
methods; the gloo unit tests pass the CPU oracle, because this module is host logic only.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of a batch of frames: ranks [0, n % world) get one extra frame."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def macroblock_row_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """Split the height/16 macroblock rows over the ranks (e.g. 270 rows of 8K over 8 GPUs -> 34,34,34,34,34,34,33,33).
    Returns pixel-row ranges [y0, y1) of the luma plane; every band is a multiple of 16 rows (possibly empty)."""
    if height % 16:
        raise ValueError("Error. height % 8 must be 0")  # chroma plane height must be a multiple of 8 (DCT.cpp:283-285)
    rows = height // 16
    return [tuple(16 * v for v in frame_range(rows, r, world)) for r in range(world)]


def slice_iyuv(iyuv: np.ndarray, w: int, h: int, y0: int, y1: int) -> np.ndarray:
    """The band [y0, y1) of an IYUV image as a contiguous IYUV image of height y1 - y0."""
    iyuv = np.asarray(iyuv, np.uint8).reshape(-1)
    Y = iyuv[: w * h].reshape(h, w)
    U = iyuv[w * h: w * h * 5 // 4].reshape(h // 2, w // 2)
    V = iyuv[w * h * 5 // 4:].reshape(h // 2, w // 2)
    return np.concatenate([Y[y0:y1].reshape(-1), U[y0 // 2: y1 // 2].reshape(-1), V[y0 // 2: y1 // 2].reshape(-1)])


def unslice_iyuv(bands: Sequence[np.ndarray], w: int, h: int, ranges: Sequence[Tuple[int, int]]) -> np.ndarray:
    out = np.empty(w * h * 3 // 2, np.uint8)
    Y = out[: w * h].reshape(h, w)
    U = out[w * h: w * h * 5 // 4].reshape(h // 2, w // 2)
    V = out[w * h * 5 // 4:].reshape(h // 2, w // 2)
    for b, (y0, y1) in zip(bands, ranges):
        bh = y1 - y0
        if bh == 0:
            continue
        b = np.asarray(b, np.uint8).reshape(-1)
        Y[y0:y1] = b[: w * bh].reshape(bh, w)
        U[y0 // 2: y1 // 2] = b[w * bh: w * bh * 5 // 4].reshape(bh // 2, w // 2)
        V[y0 // 2: y1 // 2] = b[w * bh * 5 // 4:].reshape(bh // 2, w // 2)
    return out


def parse_payload(payload: np.ndarray):
    """[(n_chunks, sizes, content)] for the three planes of a compressed payload (layout: DCT.cpp:16-73,112-173)."""
    p = np.asarray(payload, np.uint8).reshape(-1)
    psz = p[:12].view("<u4")
    planes, pos = [], 12
    for i in range(3):
        n, content = (int(v) for v in p[pos: pos + 8].view("<u4"))
        sizes = p[pos + 8: pos + 8 + n]
        planes.append((n, sizes, p[pos + 8 + n: pos + 8 + n + content]))
        pos += int(psz[i])
    return planes


def build_payload(planes) -> np.ndarray:
    """Inverse of parse_payload: planes = [(sizes, content)] * 3."""
    parts, psz = [], []
    for sizes, content in planes:
        sizes = np.asarray(sizes, np.uint8)
        content = np.asarray(content, np.uint8)
        hdr = np.array([sizes.size, content.size], "<u4").view(np.uint8)
        parts += [hdr, sizes, content]
        psz.append(8 + sizes.size + content.size)
    return np.concatenate([np.array(psz, "<u4").view(np.uint8)] + parts)


def assemble_payload(band_payloads: Sequence[np.ndarray]) -> np.ndarray:
    """One payload for the whole image from the payloads of its bands (top to bottom): per plane the chunk-size
    arrays of all bands, then the contents of all bands -- raster order of blocks is preserved because a band is
    a run of complete block rows."""
    parsed = [parse_payload(b) for b in band_payloads if b is not None and len(b)]
    planes = []
    for p in range(3):
        planes.append((np.concatenate([bp[p][1] for bp in parsed]), np.concatenate([bp[p][2] for bp in parsed])))
    return build_payload(planes)


def split_payload(payload: np.ndarray, w: int, ranges: Sequence[Tuple[int, int]]) -> List[np.ndarray]:
    """Cut a payload into per-band payloads.  Needs the prefix sum of the chunk sizes of every plane (the same scan
    the decoder does) to find each band's content range."""
    planes = parse_payload(payload)
    out = []
    for (y0, y1) in ranges:
        if y1 == y0:
            out.append(np.zeros(0, np.uint8))
            continue
        band = []
        for p, (n, sizes, content) in enumerate(planes):
            bw = (w if p == 0 else w // 2) // 8
            r0, r1 = (y0, y1) if p == 0 else (y0 // 2, y1 // 2)
            k0, k1 = (r0 // 8) * bw, (r1 // 8) * bw
            csum = np.concatenate([[0], np.cumsum(sizes, dtype=np.int64)])
            band.append((sizes[k0:k1], content[csum[k0]: csum[k1]]))
        out.append(build_payload(band))
    return out


# ------------------------------------------------------------------------------------------------
# distributed drivers (torch.distributed; backend nccl on GPUs, gloo in the CPU tests)
# ------------------------------------------------------------------------------------------------
def _gather_bytes(local: np.ndarray, dist, device) -> List[np.ndarray] | None:
    """Variable-size gather to rank 0: all-gather of the sizes, then padded gather of the bytes."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    buf[: local.size] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    if dist.get_backend() == "nccl":
        outs = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
        dist.all_gather(outs, buf)  # ~MBs over NVLink: latency bound, simpler than grouped send/recv
        got = outs if rank == 0 else None
    else:
        got = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, got, dst=0)
    if rank != 0:
        return None
    return [g[:s].cpu().numpy() for g, s in zip(got, sizes)]


def compress_image_sharded(band: np.ndarray, w: int, band_h: int, q, compress_fn: Callable, dist, device="cpu"):
    """Every rank passes ITS band (IYUV, height band_h, bands ordered by rank from top to bottom).  Returns the
    assembled payload of the whole image on rank 0 (None elsewhere)."""
    local = compress_fn(band, w, band_h, q) if band_h > 0 else np.zeros(0, np.uint8)
    parts = _gather_bytes(np.asarray(local, np.uint8), dist, device)
    if parts is None:
        return None
    return assemble_payload(parts)


def decompress_image_sharded(payload, w: int, h: int, q, decompress_fn: Callable, dist, device="cpu"):
    """Rank 0 passes the payload (others None).  Returns the decoded image on rank 0 (None elsewhere)."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    ranges = macroblock_row_bands(h, world)
    if rank == 0:
        subs = split_payload(payload, w, ranges)
        sizes = torch.tensor([s.size for s in subs], dtype=torch.int64, device=device)
    else:
        subs, sizes = None, torch.zeros(world, dtype=torch.int64, device=device)
    dist.broadcast(sizes, src=0)
    cap = max(int(sizes.max().item()), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=device)
    if dist.get_backend() == "nccl":
        # scatter is emulated with a broadcast of the padded stack (tiny), keeping to collectives NCCL always has
        stack = torch.zeros((world, cap), dtype=torch.uint8, device=device)
        if rank == 0:
            for i, s in enumerate(subs):
                stack[i, : s.size] = torch.from_numpy(s).to(device)
        dist.broadcast(stack, src=0)
        mine = stack[rank]
    else:
        src_list = None
        if rank == 0:
            src_list = []
            for s in subs:
                t = torch.zeros(cap, dtype=torch.uint8, device=device)
                t[: s.size] = torch.from_numpy(s)
                src_list.append(t)
        dist.scatter(mine, src_list, src=0)
    y0, y1 = ranges[rank]
    n = int(sizes[rank].item())
    band = decompress_fn(mine[:n].cpu().numpy(), w, y1 - y0, q) if y1 > y0 else np.zeros(0, np.uint8)
    bands = _gather_bytes(np.asarray(band, np.uint8), dist, device)
    if bands is None:
        return None
    return unslice_iyuv(bands, w, h, ranges)


# ------------------------------------------------------------------------------------------------
# device path: every band goes straight from the coding GPU into the root's buffer over NVLink
# ------------------------------------------------------------------------------------------------
class ShardGroup:
    """One rank's handle on a group of GPUs that code ONE image together (C ABI: myyuvb_dct_{compress,decompress}_shard_dev).

    Set-up (once): every rank allocates a 400-byte control block, the root allocates the payload buffer and a frame
    buffer; all of them are shared through CUDA IPC handles (``distributed``) or, for virtual ranks that live in one
    process on one device, by plain pointers (``local``).  After that a call exchanges nothing through the host: the
    ranks' streams meet on the device (kernels.cu, shard_exchange_kernel / shard_done_kernel).

    ``compress(d_iyuv)`` / ``decompress(payload_size)`` are asynchronous on the context's stream and must be issued by
    every rank of the group; ``result()`` (root) synchronises and returns the assembled size."""

    def __init__(self, ctx, rank: int, world: int, root: int, w: int, h: int, ctrl, root_out: int, out_capacity: int, root_iyuv: int,
                 owned=(), opened=()):
        from . import capi

        self.ctx, self.rank, self.world, self.root, self.w, self.h = ctx, rank, world, root, w, h
        self.ctrl, self.root_out, self.out_capacity, self.root_iyuv = list(ctrl), root_out, out_capacity, root_iyuv
        self.rows = capi.shard_rows(h, world)
        self.epoch = 0
        self._owned, self._opened = list(owned), list(opened)

    @property
    def band(self) -> Tuple[int, int]:
        return self.rows[self.rank], self.rows[self.rank + 1]

    @classmethod
    def local(cls, ctxs, w: int, h: int, root: int = 0) -> "List[ShardGroup]":
        """Virtual ranks: several contexts (streams) of one process on one device; used by the single-GPU tests."""
        from . import capi

        world = len(ctxs)
        ctrl = [c.ipc_alloc(capi.shard_ctrl_bytes())[0] for c in ctxs]
        cap = capi.compress_bound(w, h)
        out = ctxs[root].ipc_alloc(cap)[0]
        frame = ctxs[root].ipc_alloc(w * h * 3 // 2)[0]
        return [cls(c, r, world, root, w, h, ctrl, out, cap, frame, owned=([ctrl[r]] + ([out, frame] if r == root else [])))
                for r, c in enumerate(ctxs)]

    @classmethod
    def distributed(cls, ctx, dist, w: int, h: int, root: int = 0, out_capacity: int | None = None) -> "ShardGroup":
        """One process per GPU (torch.distributed initialised): IPC handles travel once through all_gather_object."""
        from . import capi

        rank, world = dist.get_rank(), dist.get_world_size()
        my_ctrl, my_handle = ctx.ipc_alloc(capi.shard_ctrl_bytes())
        mine = {"ctrl": my_handle}
        owned = [my_ctrl]
        cap = out_capacity or capi.compress_bound(w, h)
        if rank == root:
            out, mine["out"] = ctx.ipc_alloc(cap)
            frame, mine["frame"] = ctx.ipc_alloc(w * h * 3 // 2)
            owned += [out, frame]
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        ctrl, opened = [], []
        for q, rec in enumerate(everyone):
            if q == rank:
                ctrl.append(my_ctrl)
            else:
                ctrl.append(ctx.ipc_open(rec["ctrl"]))
                opened.append(ctrl[-1])
        if rank != root:
            out, frame = ctx.ipc_open(everyone[root]["out"]), ctx.ipc_open(everyone[root]["frame"])
            opened += [out, frame]
        return cls(ctx, rank, world, root, w, h, ctrl, out, cap, frame, owned=owned, opened=opened)

    def compress(self, d_iyuv, q, full_frame: bool = False) -> None:
        """d_iyuv: this rank's band as an IYUV image of its own (or the whole frame when full_frame)."""
        self.epoch += 1
        self.ctx.compress_shard_dev(d_iyuv, full_frame, self.w, self.h, q, self.rank, self.world, self.root, self.rows, self.ctrl,
                                    self.root_out, self.out_capacity, self.epoch)

    def decompress(self, payload_size: int, q, d_band_out, to_root: bool = True) -> None:
        """Decodes this rank's band of the payload in the root's buffer into d_band_out and, if to_root, into the root's frame."""
        self.epoch += 1
        self.ctx.decompress_shard_dev(self.root_out, payload_size, self.w, self.h, q, self.rank, self.world, self.root, self.rows, self.ctrl,
                                      d_band_out, self.root_iyuv if to_root else None, self.epoch)

    def result(self) -> int:
        return self.ctx.shard_result(self.ctrl[self.rank])

    def close(self) -> None:
        self.ctx.sync()
        for p in self._opened:
            self.ctx.ipc_close(p)
        for p in self._owned:
            self.ctx.ipc_free(p)
        self._opened, self._owned = [], []
