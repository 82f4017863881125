"""ctypes binding of the C ABI in include/myyuvb200.h (lib/libmyyuvb200.so).

Everything here calls into the CUDA library; nothing is computed in Python.  Loading fails loudly when
the library has not been built (``python -m yuv-manipulations-2_b200.build`` / ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import pathlib

import numpy as np

_PKG = pathlib.Path(__file__).resolve().parent
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)

# error codes of include/myyuvb200.h
OK, ERR_CUDA, ERR_ARG, ERR_QUALITY, ERR_WIDTH, ERR_HEIGHT, ERR_CAPACITY, ERR_DCTYUV_SIZE, ERR_PLANE_SIZE, ERR_HUFFMAN, \
    ERR_EVEN, ERR_TOO_LARGE, ERR_SHARD_TIMEOUT, ERR_BOUNDS = range(14)

EXPORTS = [
    "myyuvb_ctx_create", "myyuvb_ctx_destroy", "myyuvb_last_error", "myyuvb_sync", "myyuvb_stream",
    "myyuvb_compress_bound", "myyuvb_xrgb_to_iyuv", "myyuvb_dct_compress", "myyuvb_dct_decompress",
    "myyuvb_xrgb_to_iyuv_batch_dev", "myyuvb_dct_compress_batch_dev", "myyuvb_dct_decompress_batch_dev",
    "myyuvb_xrgb_dct_compress_batch_dev", "myyuvb_bgr24_to_iyuv", "myyuvb_bgr24_to_iyuv_batch_dev",
    "myyuvb_batch_status", "myyuvb_dct_compress_batch_host", "myyuvb_dct_decompress_batch_host",
    "myyuvb_host_alloc", "myyuvb_host_free", "myyuvb_launch_count", "myyuvb_last_kernel_ms",
    "myyuvb_dct_compress_begin", "myyuvb_dct_compress_fetch", "myyuvb_phase_clocks",
    "myyuvb_shard_ctrl_bytes", "myyuvb_shard_rows", "myyuvb_dct_compress_shard_dev", "myyuvb_dct_decompress_shard_dev",
    "myyuvb_set_encoder_mode", "myyuvb_shard_result", "myyuvb_ipc_alloc",
    "myyuvb_iyuv_planes", "myyuvb_get_pixels_dev", "myyuvb_iyuv_to_rgba_batch_dev", "myyuvb_dct_decompress_to_rgba_batch_dev", "myyuvb_ipc_open", "myyuvb_ipc_close", "myyuvb_ipc_free",
]


class MyyuvError(RuntimeError):
    """Raised with the message the reference would have thrown (std::runtime_error) and the C-ABI code."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def library_path() -> pathlib.Path:
    import os

    variant = os.environ.get("MYYUVB_LIB_VARIANT", "")  # "clk": the profiling build with per-phase clock counters
    return _PKG / "lib" / ("libmyyuvb200.so" if not variant else f"libmyyuvb200_{variant}.so")


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise ImportError(f"{path} is missing: build the CUDA library first (python __graft_entry__.py build). "
                          "There is no CPU fallback.")
    L = C.CDLL(str(path))
    L.myyuvb_last_error.restype = C.c_char_p
    L.myyuvb_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.myyuvb_ctx_destroy.argtypes = [C.c_void_p]
    L.myyuvb_ctx_destroy.restype = None
    L.myyuvb_sync.argtypes = [C.c_void_p]
    L.myyuvb_stream.argtypes = [C.c_void_p]
    L.myyuvb_stream.restype = C.c_void_p
    L.myyuvb_compress_bound.argtypes = [C.c_uint32, C.c_uint32]
    L.myyuvb_compress_bound.restype = C.c_uint64
    L.myyuvb_xrgb_to_iyuv.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    L.myyuvb_bgr24_to_iyuv.argtypes = L.myyuvb_xrgb_to_iyuv.argtypes
    L.myyuvb_dct_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_void_p, C.c_uint64,
                                      C.POINTER(C.c_uint32)]
    L.myyuvb_dct_compress_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.POINTER(C.c_uint32)]
    L.myyuvb_dct_compress_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.myyuvb_dct_decompress.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, C.c_void_p]
    L.myyuvb_xrgb_to_iyuv_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p]
    L.myyuvb_bgr24_to_iyuv_batch_dev.argtypes = L.myyuvb_xrgb_to_iyuv_batch_dev.argtypes
    L.myyuvb_dct_compress_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, C.c_void_p,
                                                C.c_uint64, C.c_void_p]
    L.myyuvb_dct_decompress_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32,
                                                  C.c_void_p]
    L.myyuvb_xrgb_dct_compress_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, _u8p, C.c_uint32,
                                                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.myyuvb_batch_status.argtypes = [C.c_void_p]
    L.myyuvb_dct_compress_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, C.c_void_p,
                                                 C.c_uint64, C.c_void_p]
    L.myyuvb_dct_decompress_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32,
                                                   C.c_void_p]
    L.myyuvb_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.myyuvb_host_free.argtypes = [C.c_void_p]
    L.myyuvb_host_free.restype = None
    L.myyuvb_launch_count.restype = C.c_uint64
    L.myyuvb_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.myyuvb_phase_clocks.argtypes = [C.c_void_p, C.c_int]
    L.myyuvb_phase_clocks.restype = None
    L.myyuvb_set_encoder_mode.argtypes = [C.c_void_p, C.c_int]
    L.myyuvb_iyuv_planes.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.myyuvb_get_pixels_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    L.myyuvb_iyuv_to_rgba_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    L.myyuvb_dct_decompress_to_rgba_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32,
                                                          C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    L.myyuvb_shard_ctrl_bytes.restype = C.c_uint64
    L.myyuvb_shard_rows.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    L.myyuvb_dct_compress_shard_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, C.c_uint32,
                                                C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.c_void_p, C.c_uint64, C.c_uint32]
    L.myyuvb_dct_decompress_shard_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, C.c_uint32,
                                                  C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_uint32]
    L.myyuvb_shard_result.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.myyuvb_ipc_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p]
    L.myyuvb_ipc_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.myyuvb_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
    L.myyuvb_ipc_free.argtypes = [C.c_void_p, C.c_void_p]
    _lib = L
    return L


def _check(rc: int):
    if rc:
        raise MyyuvError(rc, lib().myyuvb_last_error().decode())


def _q(q) -> np.ndarray:
    qi = np.asarray(q).astype(np.int64).reshape(-1)
    if qi.size != 3:
        # compress_map lambda, myyuv_yuv.cpp:134-136
        raise MyyuvError(ERR_ARG, "Error compression: incorrect parameters count. 3 parameters required")
    if ((qi < 1) | (qi > 100)).any():
        # checked before the cast to uint8, which would fold e.g. 306 onto 50 (DCT.cpp:378-382 tests the uint8 the caller stored)
        raise MyyuvError(ERR_QUALITY, "Level of quality must be between 1 and 100")
    return np.ascontiguousarray(qi.astype(np.uint8))


def compress_bound(w: int, h: int) -> int:
    return int(lib().myyuvb_compress_bound(w, h))


def phase_clocks(reset: bool = True) -> np.ndarray:
    """[2][12] clock sums per phase (compress, decompress); zeros unless MYYUVB_LIB_VARIANT=clk."""
    out = np.zeros(24, np.uint64)
    lib().myyuvb_phase_clocks(out.ctypes.data, int(reset))
    return out.reshape(2, 12)


def iyuv_planes(addr: int, w: int, h: int):
    """[(address, width, height)] of the Y, U, V planes of an IYUV frame at `addr` (host or device)."""
    pl = (C.c_void_p * 3)()
    ws, hs = (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
    _check(lib().myyuvb_iyuv_planes(C.c_void_p(addr), w, h, pl, ws, hs))
    return [(int(pl[i]), int(ws[i]), int(hs[i])) for i in range(3)]


def shard_rows(height: int, world: int):
    """Balanced split of the height/16 macroblock rows over the ranks: luma row boundaries, world + 1 entries."""
    rows = (C.c_uint32 * (world + 1))()
    _check(lib().myyuvb_shard_rows(height, world, rows))
    return list(rows)


def shard_ctrl_bytes() -> int:
    return int(lib().myyuvb_shard_ctrl_bytes())


def launch_count() -> int:
    return int(lib().myyuvb_launch_count())


class PinnedBuffer:
    """Page-locked host memory from the library (cudaHostAlloc), viewed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        _check(lib().myyuvb_host_alloc(nbytes, C.byref(p)))
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(C.cast(p, _u8p), shape=(max(nbytes, 1),))[:nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            lib().myyuvb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _ptr(x) -> int:
    """Address of a numpy array (host) or a torch tensor (host or device)."""
    if isinstance(x, np.ndarray):
        assert x.flags.c_contiguous
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return x.data_ptr()
    if isinstance(x, int):
        return x
    raise TypeError(type(x))


class Context:
    """One CUDA device + stream + reusable scratch (myyuvb_ctx).  Not thread-safe; use one per thread."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        _check(lib().myyuvb_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().myyuvb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        _check(lib().myyuvb_sync(self._h))

    def set_encoder_mode(self, mode: int) -> None:
        """0: automatic, 1: queue blocks with more than 8 symbols, 2: code up to 15 symbols in place (same bytes either way)."""
        _check(lib().myyuvb_set_encoder_mode(self._h, mode))

    @property
    def stream(self) -> int:
        return int(lib().myyuvb_stream(self._h) or 0)

    # ---- host-pointer single image calls (numpy in, numpy out) ----
    def xrgb_to_iyuv(self, bgrx: np.ndarray, w: int, h: int, bottom_up: bool = True) -> np.ndarray:
        bgrx = np.ascontiguousarray(bgrx, dtype=np.uint8).reshape(-1)
        if bgrx.size != w * h * 4:
            raise ValueError("bgrx must hold width*height*4 bytes")
        out = np.empty(w * h * 3 // 2, np.uint8)
        _check(lib().myyuvb_xrgb_to_iyuv(self._h, bgrx.ctypes.data, w, h, int(bottom_up), out.ctypes.data))
        return out

    def bgr24_to_iyuv(self, bgr: np.ndarray, w: int, h: int, bottom_up: bool = True) -> np.ndarray:
        """24-bit BMP pixel rows (B,G,R triplets, no row padding)."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8).reshape(-1)
        if bgr.size != w * h * 3:
            raise ValueError("bgr must hold width*height*3 bytes")
        out = np.empty(w * h * 3 // 2, np.uint8)
        _check(lib().myyuvb_bgr24_to_iyuv(self._h, bgr.ctypes.data, w, h, int(bottom_up), out.ctypes.data))
        return out

    def compress(self, iyuv: np.ndarray, w: int, h: int, q, capacity: int | None = None) -> np.ndarray:
        iyuv = np.ascontiguousarray(iyuv, dtype=np.uint8).reshape(-1)
        if iyuv.size != w * h * 3 // 2:
            raise ValueError("iyuv must hold width*height*3/2 bytes")
        qa = _q(q)
        n = C.c_uint32(0)
        if capacity is None:  # two steps: the size first, then an exact-size buffer (what the class API does)
            _check(lib().myyuvb_dct_compress_begin(self._h, iyuv.ctypes.data, w, h, qa.ctypes.data_as(_u8p), C.byref(n)))
            out = np.empty(max(n.value, 1), np.uint8)
            _check(lib().myyuvb_dct_compress_fetch(self._h, out.ctypes.data, n.value))
            return out[: n.value]
        out = np.empty(max(capacity, 1), np.uint8)
        _check(lib().myyuvb_dct_compress(self._h, iyuv.ctypes.data, w, h, qa.ctypes.data_as(_u8p), out.ctypes.data, capacity, C.byref(n)))
        return out[: n.value].copy()

    def decompress(self, payload: np.ndarray, w: int, h: int, q) -> np.ndarray:
        payload = np.ascontiguousarray(payload, dtype=np.uint8).reshape(-1)
        qa = _q(q)
        out = np.empty(w * h * 3 // 2, np.uint8)
        _check(lib().myyuvb_dct_decompress(self._h, payload.ctypes.data if payload.size else None, payload.size, w, h,
                                           qa.ctypes.data_as(_u8p), out.ctypes.data))
        return out

    # ---- host-pointer batch calls (pipelined H2D / kernels / D2H) ----
    def compress_batch_host(self, iyuv, w: int, h: int, q, n_frames: int, out, offsets: np.ndarray) -> None:
        qa = _q(q)
        assert offsets.dtype == np.uint64 and offsets.size >= n_frames + 1
        nbytes = out.nbytes if isinstance(out, np.ndarray) else out.numel()
        _check(lib().myyuvb_dct_compress_batch_host(self._h, _ptr(iyuv), w, h, qa.ctypes.data_as(_u8p), n_frames, _ptr(out), nbytes,
                                                    offsets.ctypes.data))

    def decompress_batch_host(self, payloads, offsets: np.ndarray, w: int, h: int, q, n_frames: int, out) -> None:
        qa = _q(q)
        assert offsets.dtype == np.uint64 and offsets.size >= n_frames + 1
        _check(lib().myyuvb_dct_decompress_batch_host(self._h, _ptr(payloads), offsets.ctypes.data, w, h, qa.ctypes.data_as(_u8p),
                                                      n_frames, _ptr(out)))

    # ---- device-pointer batch calls (torch CUDA tensors or raw device addresses; asynchronous) ----
    def xrgb_to_iyuv_batch_dev(self, d_bgrx, w: int, h: int, bottom_up: bool, n_frames: int, d_iyuv) -> None:
        _check(lib().myyuvb_xrgb_to_iyuv_batch_dev(self._h, _ptr(d_bgrx), w, h, int(bottom_up), n_frames, _ptr(d_iyuv)))

    def bgr24_to_iyuv_batch_dev(self, d_bgr, w: int, h: int, bottom_up: bool, n_frames: int, d_iyuv) -> None:
        _check(lib().myyuvb_bgr24_to_iyuv_batch_dev(self._h, _ptr(d_bgr), w, h, int(bottom_up), n_frames, _ptr(d_iyuv)))

    def compress_batch_dev(self, d_iyuv, w: int, h: int, q, n_frames: int, d_out, out_capacity: int, d_offsets) -> None:
        qa = _q(q)
        _check(lib().myyuvb_dct_compress_batch_dev(self._h, _ptr(d_iyuv), w, h, qa.ctypes.data_as(_u8p), n_frames, _ptr(d_out),
                                                   out_capacity, _ptr(d_offsets)))

    def decompress_batch_dev(self, d_payloads, d_offsets, w: int, h: int, q, n_frames: int, d_iyuv) -> None:
        qa = _q(q)
        _check(lib().myyuvb_dct_decompress_batch_dev(self._h, _ptr(d_payloads), _ptr(d_offsets), w, h, qa.ctypes.data_as(_u8p),
                                                     n_frames, _ptr(d_iyuv)))

    def xrgb_compress_batch_dev(self, d_bgrx, w: int, h: int, bottom_up: bool, q, n_frames: int, d_out, out_capacity: int,
                                d_offsets, d_iyuv=None, chunk_frames: int = 0) -> None:
        """YUV(bmp, IYUV).compress(DCT, q) for a device-resident batch of XRGB frames (chunked so that the IYUV
        intermediate is read back from L2); d_iyuv optionally receives the IYUV frames too."""
        qa = _q(q)
        _check(lib().myyuvb_xrgb_dct_compress_batch_dev(self._h, _ptr(d_bgrx), w, h, int(bottom_up), qa.ctypes.data_as(_u8p), n_frames,
                                                        chunk_frames, _ptr(d_iyuv) if d_iyuv is not None else None, _ptr(d_out),
                                                        out_capacity, _ptr(d_offsets)))

    # ---- consumers of decoded frames on the device: getPixel, display RGB ----
    def get_pixels_dev(self, d_iyuv, w: int, h: int, n: int, d_xy, d_out) -> None:
        _check(lib().myyuvb_get_pixels_dev(self._h, _ptr(d_iyuv), w, h, n, _ptr(d_xy), _ptr(d_out)))

    def iyuv_to_rgba_batch_dev(self, d_iyuv, w: int, h: int, n_frames: int, d_rgba, flip_rows: bool = False) -> None:
        _check(lib().myyuvb_iyuv_to_rgba_batch_dev(self._h, _ptr(d_iyuv), w, h, n_frames, int(flip_rows), _ptr(d_rgba)))

    def decompress_to_rgba_batch_dev(self, d_payloads, d_offsets, w: int, h: int, q, n_frames: int, d_rgba, d_iyuv=None,
                                     chunk_frames: int = 0, flip_rows: bool = False) -> None:
        qa = _q(q)
        _check(lib().myyuvb_dct_decompress_to_rgba_batch_dev(self._h, _ptr(d_payloads), _ptr(d_offsets), w, h, qa.ctypes.data_as(_u8p), n_frames,
                                                             chunk_frames, int(flip_rows), _ptr(d_iyuv) if d_iyuv is not None else None,
                                                             _ptr(d_rgba)))

    # ---- one image sharded over the GPUs of a box (see sharding.ShardGroup for the set-up) ----
    def ipc_alloc(self, nbytes: int):
        """(device address, 64-byte IPC handle) of zeroed device memory other processes can map."""
        p = C.c_void_p()
        h = (C.c_uint8 * 64)()
        _check(lib().myyuvb_ipc_alloc(self._h, nbytes, C.byref(p), h))
        return int(p.value), bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        _check(lib().myyuvb_ipc_open(self._h, buf, C.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int) -> None:
        _check(lib().myyuvb_ipc_close(self._h, C.c_void_p(ptr)))

    def ipc_free(self, ptr: int) -> None:
        _check(lib().myyuvb_ipc_free(self._h, C.c_void_p(ptr)))

    def compress_shard_dev(self, d_iyuv, full_frame: bool, w: int, h: int, q, rank: int, world: int, root: int, rows, ctrl,
                           root_out: int, out_capacity: int, epoch: int) -> None:
        qa = _q(q)
        ra = (C.c_uint32 * (world + 1))(*rows)
        ca = (C.c_void_p * world)(*ctrl)
        _check(lib().myyuvb_dct_compress_shard_dev(self._h, _ptr(d_iyuv), int(full_frame), w, h, qa.ctypes.data_as(_u8p), rank, world, root,
                                                   ra, ca, C.c_void_p(root_out), out_capacity, epoch))

    def decompress_shard_dev(self, root_payload: int, payload_size: int, w: int, h: int, q, rank: int, world: int, root: int, rows, ctrl,
                             d_band_out, root_iyuv: int | None, epoch: int) -> None:
        qa = _q(q)
        ra = (C.c_uint32 * (world + 1))(*rows)
        ca = (C.c_void_p * world)(*ctrl)
        _check(lib().myyuvb_dct_decompress_shard_dev(self._h, C.c_void_p(root_payload), payload_size, w, h, qa.ctypes.data_as(_u8p), rank,
                                                     world, root, ra, ca, _ptr(d_band_out), C.c_void_p(root_iyuv) if root_iyuv else None, epoch))

    def shard_result(self, ctrl_local: int) -> int:
        n = C.c_uint64(0)
        _check(lib().myyuvb_shard_result(self._h, C.c_void_p(ctrl_local), C.byref(n)))
        return int(n.value)

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib().myyuvb_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def batch_status(self) -> None:
        """Synchronise and raise the first data-dependent error of the batch calls issued so far."""
        _check(lib().myyuvb_batch_status(self._h))


_default_ctx: Context | None = None


def default_context() -> Context:
    """Lazily created context on device 0 (what the class API uses, like the reference's free functions)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx
