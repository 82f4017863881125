"""oracle -- TEST INFRASTRUCTURE ONLY (ctypes front-ends of the CPU checkers).

* ``Oracle``    : oracle/liboracle.so, the plain-C restatement in myyuv_oracle.c ("port").
* ``Reference`` : oracle/_ref/{omp,serial}/librefshim.so, the UNMODIFIED reference compiled from
                  /root/reference by oracle/Makefile, driven through ref_shim.cpp ("reference").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package never does: it must fail loudly without its CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
GOLDEN_DIR = REF_DIR / "golden"

_u8p = C.POINTER(C.c_uint8)
_i16p = C.POINTER(C.c_int16)


def _p8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data_as(_u8p)


def _p16(a: np.ndarray):
    assert a.dtype == np.int16 and a.flags.c_contiguous
    return a.ctypes.data_as(_i16p)


def build(ref_root: str = "/root/reference") -> None:
    """Compile liboracle.so and, when the reference checkout exists, oracle/_ref (see Makefile)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle", "ref", f"REF={ref_root}"], check=True)


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"oracle error code {code}")
        self.code = code


class Oracle:
    """The C restatement.  Method names follow the reference's operations."""

    ERR_QUALITY, ERR_WIDTH, ERR_HEIGHT, ERR_CAPACITY, ERR_DCTYUV_SIZE, ERR_PLANE_SIZE, ERR_HUFF_CODE, ERR_CHUNK = range(1, 9)

    def __init__(self, threads: int | None = None):
        path = HERE / "liboracle.so"
        if not path.exists():
            build()
        if threads is not None:
            os.environ["OMP_NUM_THREADS"] = str(threads)
        self.lib = C.CDLL(str(path))
        L = self.lib
        L.ora_bgrx_to_iyuv.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_int, _u8p]
        L.ora_bgrx_to_iyuv.restype = None
        L.ora_bgr_to_iyuv_pb.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, _u8p]
        L.ora_bgr_to_iyuv_pb.restype = None
        L.ora_qtable.argtypes = [C.c_uint8, C.c_int, C.POINTER(C.c_float)]
        L.ora_qtable.restype = None
        L.ora_compress_bound.argtypes = [C.c_uint32, C.c_uint32]
        L.ora_compress_bound.restype = C.c_uint32
        L.ora_compress.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p, _u8p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.ora_decompress.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p]
        L.ora_plane_coefs.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint8, C.c_int, _i16p]
        L.ora_payload_coefs.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, _i16p]
        L.ora_huff_encode_blocks.argtypes = [_i16p, C.c_uint32, _u8p, _u8p]
        L.ora_huff_encode_blocks.restype = None
        L.ora_huff_decode_blocks.argtypes = [_u8p, _u8p, C.c_uint32, _i16p]

    def bgrx_to_iyuv(self, bgrx: np.ndarray, w: int, h: int, bottom_up: bool = True) -> np.ndarray:
        bgrx = np.ascontiguousarray(bgrx, dtype=np.uint8).reshape(-1)
        assert bgrx.size == w * h * 4
        out = np.empty(w * h * 3 // 2, np.uint8)
        self.lib.ora_bgrx_to_iyuv(_p8(bgrx), w, h, int(bottom_up), _p8(out))
        return out

    def bgr24_to_iyuv(self, bgr: np.ndarray, w: int, h: int, bottom_up: bool = True) -> np.ndarray:
        """24-bit BMP pixel rows (B,G,R; width % 4 == 0 so rows carry no padding, myyuv_bmp.cpp:130)."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8).reshape(-1)
        assert bgr.size == w * h * 3
        out = np.empty(w * h * 3 // 2, np.uint8)
        self.lib.ora_bgr_to_iyuv_pb(_p8(bgr), w, h, int(bottom_up), 3, _p8(out))
        return out

    def qtable(self, q: int, chroma: bool) -> np.ndarray:
        qt = np.empty(64, np.float32)
        self.lib.ora_qtable(q, int(chroma), qt.ctypes.data_as(C.POINTER(C.c_float)))
        return qt

    def compress(self, iyuv: np.ndarray, w: int, h: int, q) -> np.ndarray:
        iyuv = np.ascontiguousarray(iyuv, dtype=np.uint8).reshape(-1)
        assert iyuv.size == w * h * 3 // 2
        qa = np.asarray(q, np.uint8)
        cap = int(self.lib.ora_compress_bound(w, h))
        out = np.empty(cap, np.uint8)
        n = C.c_uint32(0)
        rc = self.lib.ora_compress(_p8(iyuv), w, h, _p8(qa), _p8(out), cap, C.byref(n))
        if rc:
            raise OracleError(rc)
        return out[: n.value].copy()

    def decompress(self, payload: np.ndarray, w: int, h: int, q) -> np.ndarray:
        payload = np.ascontiguousarray(payload, dtype=np.uint8).reshape(-1)
        qa = np.asarray(q, np.uint8)
        out = np.empty(w * h * 3 // 2, np.uint8)
        rc = self.lib.ora_decompress(_p8(payload), payload.size, w, h, _p8(qa), _p8(out))
        if rc:
            raise OracleError(rc)
        return out

    def plane_coefs(self, plane: np.ndarray, w: int, h: int, q: int, chroma: bool) -> np.ndarray:
        plane = np.ascontiguousarray(plane, dtype=np.uint8).reshape(-1)
        out = np.empty((w * h // 64, 64), np.int16)
        rc = self.lib.ora_plane_coefs(_p8(plane), w, h, q, int(chroma), _p16(out))
        if rc:
            raise OracleError(rc)
        return out

    def payload_coefs(self, payload: np.ndarray, w: int, h: int) -> np.ndarray:
        payload = np.ascontiguousarray(payload, dtype=np.uint8).reshape(-1)
        out = np.empty((w * h // 64 * 3 // 2, 64), np.int16)
        rc = self.lib.ora_payload_coefs(_p8(payload), payload.size, w, h, _p16(out))
        if rc:
            raise OracleError(rc)
        return out

    def huff_encode_blocks(self, coefs: np.ndarray):
        coefs = np.ascontiguousarray(coefs, dtype=np.int16).reshape(-1, 64)
        n = coefs.shape[0]
        out = np.empty(n * 256, np.uint8)
        sizes = np.empty(n, np.uint8)
        self.lib.ora_huff_encode_blocks(_p16(coefs), n, _p8(out), _p8(sizes))
        return out[: int(sizes.sum(dtype=np.int64))].copy(), sizes

    def huff_decode_blocks(self, chunks: np.ndarray, sizes: np.ndarray) -> np.ndarray:
        chunks = np.ascontiguousarray(chunks, dtype=np.uint8)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint8)
        out = np.empty((sizes.size, 64), np.int16)
        rc = self.lib.ora_huff_decode_blocks(_p8(chunks), _p8(sizes), sizes.size, _p16(out))
        if rc:
            raise OracleError(rc)
        return out


class ReferenceUnavailable(RuntimeError):
    pass


class Reference:
    """The unmodified reference library behind ref_shim.cpp.  variant: 'omp' (README's
    -DMYYUV_USE_OPENMP=ON build) or 'serial' (default CMake build)."""

    def __init__(self, variant: str = "omp"):
        path = REF_DIR / variant / "librefshim.so"
        if not path.exists():
            raise ReferenceUnavailable(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.variant = variant
        self.lib = C.CDLL(str(path))
        L = self.lib
        L.refshim_last_error.restype = C.c_char_p
        L.refshim_threads.restype = C.c_int
        dp = C.POINTER(C.c_double)
        L.refshim_bgrx_to_iyuv.argtypes = [_u8p, C.c_int32, C.c_int32, _u8p, dp]
        L.refshim_bmp_to_iyuv.argtypes = [_u8p, C.c_int32, C.c_int32, C.c_uint32, _u8p, dp]
        L.refshim_compress.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p, C.c_uint32, _u8p, C.c_uint32,
                                       C.POINTER(C.c_uint32), dp]
        L.refshim_decompress.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p, _u8p, dp]
        L.refshim_huffman_encode.argtypes = [_i16p, C.c_uint32, _u8p, _u8p]
        L.refshim_huffman_decode.argtypes = [_u8p, _u8p, C.c_uint32, _i16p]
        self.last_seconds = 0.0

    @property
    def threads(self) -> int:
        return int(self.lib.refshim_threads())

    def _check(self, rc: int):
        if rc:
            raise RuntimeError(self.lib.refshim_last_error().decode())

    def bgrx_to_iyuv(self, bgrx: np.ndarray, w: int, h: int, bottom_up: bool = True) -> np.ndarray:
        bgrx = np.ascontiguousarray(bgrx, dtype=np.uint8).reshape(-1)
        out = np.empty(w * h * 3 // 2, np.uint8)
        sec = C.c_double(0)
        self._check(self.lib.refshim_bgrx_to_iyuv(_p8(bgrx), w, h if bottom_up else -h, _p8(out), C.byref(sec)))
        self.last_seconds = sec.value
        return out

    def bgr24_to_iyuv(self, bgr: np.ndarray, w: int, h: int, bottom_up: bool = True) -> np.ndarray:
        """The reference's converter on a 24-bit BMP (its 32-bit assert is compiled out by -DNDEBUG, oracle/Makefile)."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8).reshape(-1)
        assert bgr.size == w * h * 3
        out = np.empty(w * h * 3 // 2, np.uint8)
        sec = C.c_double(0)
        self._check(self.lib.refshim_bmp_to_iyuv(_p8(bgr), w, h if bottom_up else -h, 24, _p8(out), C.byref(sec)))
        self.last_seconds = sec.value
        return out

    def compress(self, iyuv: np.ndarray, w: int, h: int, q) -> np.ndarray:
        iyuv = np.ascontiguousarray(iyuv, dtype=np.uint8).reshape(-1)
        qa = np.asarray(q, np.uint8)
        cap = 12 + 24 + (w * h // 64 * 3 // 2) * 256
        out = np.empty(cap, np.uint8)
        n = C.c_uint32(0)
        sec = C.c_double(0)
        self._check(self.lib.refshim_compress(_p8(iyuv), w, h, _p8(qa), qa.size, _p8(out), cap, C.byref(n), C.byref(sec)))
        self.last_seconds = sec.value
        return out[: n.value].copy()

    def decompress(self, payload: np.ndarray, w: int, h: int, q) -> np.ndarray:
        payload = np.ascontiguousarray(payload, dtype=np.uint8).reshape(-1)
        qa = np.asarray(q, np.uint8)
        out = np.empty(w * h * 3 // 2, np.uint8)
        sec = C.c_double(0)
        self._check(self.lib.refshim_decompress(_p8(payload), payload.size, w, h, _p8(qa), _p8(out), C.byref(sec)))
        self.last_seconds = sec.value
        return out

    def huff_encode_blocks(self, coefs: np.ndarray):
        coefs = np.ascontiguousarray(coefs, dtype=np.int16).reshape(-1, 64)
        n = coefs.shape[0]
        out = np.empty(n * 256, np.uint8)
        sizes = np.empty(n, np.uint8)
        self._check(self.lib.refshim_huffman_encode(_p16(coefs), n, _p8(out), _p8(sizes)))
        return out[: int(sizes.sum(dtype=np.int64))].copy(), sizes

    def huff_decode_blocks(self, chunks: np.ndarray, sizes: np.ndarray) -> np.ndarray:
        chunks = np.ascontiguousarray(chunks, dtype=np.uint8)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint8)
        out = np.empty((sizes.size, 64), np.int16)
        self._check(self.lib.refshim_huffman_decode(_p8(chunks), _p8(sizes), sizes.size, _p16(out)))
        return out


def have_reference() -> bool:
    return (REF_DIR / "omp" / "librefshim.so").exists()


# ---- .myyuv / .bmp file helpers for the tests (format: myyuv_yuv.hpp:13-29, myyuv_bmp.hpp:12-43) ----
import struct

YUV_HDR = struct.Struct("<2sIIHIIIII32s")  # type, fourcc, data_size, compression, params_size, params_pos, w, h, data_pos, unused


def read_myyuv(path):
    raw = pathlib.Path(path).read_bytes()
    typ, fourcc, data_size, comp, psz, ppos, w, h, dpos, _ = YUV_HDR.unpack_from(raw, 0)
    assert typ == b"YU" and fourcc == 0x56555949
    params = np.frombuffer(raw, np.uint8, psz, ppos).copy() if psz else np.zeros(0, np.uint8)
    if comp == 0:
        data_size = w * h * 3 // 2
    data = np.frombuffer(raw, np.uint8, data_size, dpos).copy()
    return dict(w=w, h=h, compression=comp, params=params, data=data)


def read_bmp32(path):
    raw = pathlib.Path(path).read_bytes()
    assert raw[:2] == b"BM"
    data_pos = struct.unpack_from("<I", raw, 10)[0]
    w, h = struct.unpack_from("<ii", raw, 18)
    bpp = struct.unpack_from("<H", raw, 28)[0]
    assert bpp == 32
    px = np.frombuffer(raw, np.uint8, abs(w) * abs(h) * 4, data_pos).copy()
    return dict(w=abs(w), h=abs(h), bottom_up=h > 0, data=px)
