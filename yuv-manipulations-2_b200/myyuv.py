"""Python mirror of the reference's class API for the hot path (myyuv_lib/myyuv_bmp.hpp, myyuv_yuv.hpp).

Same names, argument meaning and error behaviour as ``myyuv::BMP`` / ``myyuv::YUV`` so that the parity tests
read like tests of the reference: ``YUV(bmp, YUV.FourccFormats.IYUV)``, ``yuv.compress(YUV.Compressions.DCT,
params)``, ``yuv.decompress()``, ``yuv.dump(path)``, and the three public registries
``YUV.bmp_to_yuv_map`` / ``YUV.compress_map`` / ``YUV.decompress_map`` (myyuv_yuv.hpp:106,111,116) whose entries
call the CUDA library through the C ABI.  Host-side only: file I/O, headers, dispatch, checks.  The shipped
drop-in for C++ users is lib/libmyyuv_lib.so (csrc/myyuv_yuv.cpp); this module is the same thing for Python.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from . import capi

_YUV_HDR = struct.Struct("<2sIIHIIIII32s")        # myyuv_yuv.hpp:13-29, packed, 64 bytes
_BMP_HDR = struct.Struct("<2sIHHIIiiHHIIiiII")    # myyuv_bmp.hpp:12-31, packed, 54 bytes
_BMP_COLOR = struct.Struct("<IIIII16I")           # myyuv_bmp.hpp:36-43, 84 bytes
assert _YUV_HDR.size == 64 and _BMP_HDR.size == 54 and _BMP_COLOR.size == 84


@dataclass
class YUVHeader:
    type: bytes = b"YU"
    fourcc_format: int = 0
    data_size: int = 0
    compression: int = 0
    compression_params_size: int = 0
    compression_params_pos: int = 0
    width: int = 0
    height: int = 0
    data_pos: int = 0
    unused: bytes = bytes(32)

    def pack(self) -> bytes:
        return _YUV_HDR.pack(self.type, self.fourcc_format, self.data_size, self.compression, self.compression_params_size,
                             self.compression_params_pos, self.width, self.height, self.data_pos, self.unused)

    @classmethod
    def unpack(cls, raw: bytes) -> "YUVHeader":
        return cls(*_YUV_HDR.unpack(raw[:64].ljust(64, b"\0")))

    def copy(self) -> "YUVHeader":
        return YUVHeader(**self.__dict__)


@dataclass
class BMPHeader:
    type: bytes = b"BM"
    file_size: int = 0
    reserved1: int = 0
    reserved2: int = 0
    data_pos: int = 0
    header_size: int = 0
    width: int = 0
    height: int = 0
    planes: int = 0
    bit_count: int = 0
    compression: int = 0
    size_image_for_compression: int = 0
    x_pixels_per_meter: int = 0
    y_pixels_per_meter: int = 0
    colors_used: int = 0
    colors_important: int = 0

    def pack(self) -> bytes:
        return _BMP_HDR.pack(*[getattr(self, f) for f in self.__dataclass_fields__])

    @classmethod
    def unpack(cls, raw: bytes) -> "BMPHeader":
        return cls(*_BMP_HDR.unpack(raw[:54].ljust(54, b"\0")))


@dataclass
class BMPColorHeader:
    red_mask: int = 0x00FF0000
    green_mask: int = 0x0000FF00
    blue_mask: int = 0x000000FF
    alpha_mask: int = 0xFF000000
    color_space: int = 0x73524742
    unused: tuple = field(default_factory=lambda: (0,) * 16)

    def pack(self) -> bytes:
        return _BMP_COLOR.pack(self.red_mask, self.green_mask, self.blue_mask, self.alpha_mask, self.color_space, *self.unused)

    @classmethod
    def unpack(cls, raw: bytes) -> "BMPColorHeader":
        v = _BMP_COLOR.unpack(raw[:84].ljust(84, b"\0"))
        return cls(v[0], v[1], v[2], v[3], v[4], tuple(v[5:]))


class BMP:
    """myyuv::BMP (myyuv_bmp.hpp:52-157): 54-byte header (+84-byte colour header for 32 bpp) + pixel rows."""

    def __init__(self, path: str | None = None):
        self.header = BMPHeader()
        self.color_header = BMPColorHeader()
        self.data: np.ndarray | None = None
        if path is not None:
            self.load(path)

    def trueWidth(self) -> int:
        return abs(self.header.width)

    def trueHeight(self) -> int:
        return abs(self.header.height)

    def imageSize(self) -> int:
        return (self.trueWidth() * self.trueHeight() * self.header.bit_count // 8) & 0xFFFFFFFF

    def isValidHeader(self) -> bool:  # myyuv_bmp.cpp:127-139
        h, c = self.header, self.color_header
        return (h.type == b"BM" and h.width % 4 == 0 and h.bit_count > 0 and h.header_size > 0 and h.compression in (0, 3)
                and h.colors_used == 0 and h.colors_important == 0 and c.red_mask == 0x00FF0000 and c.green_mask == 0x0000FF00
                and c.blue_mask == 0x000000FF and c.alpha_mask in (0xFF000000, 0) and c.color_space == 0x73524742)

    def isValid(self) -> bool:
        return self.data is not None and self.isValidHeader()

    def load(self, path: str) -> None:  # myyuv_bmp.cpp:141-168
        try:
            raw = open(path, "rb").read()
        except OSError:
            raise RuntimeError("Error opening file to read " + path)
        res = BMP()
        res.header = BMPHeader.unpack(raw)
        if res.header.bit_count == 32:
            res.color_header = BMPColorHeader.unpack(raw[54:])
        pos = res.header.data_pos
        res.header.data_pos = 54 + 84 if res.header.bit_count == 32 else 54
        size = res.imageSize()
        res.header.file_size = res.header.data_pos + size
        if not res.isValidHeader():
            raise RuntimeError("Error bad header " + path)
        res.data = np.frombuffer(raw[pos:pos + size].ljust(size, b"\0"), np.uint8).copy()
        self.header, self.color_header, self.data = res.header, res.color_header, res.data

    def dump(self, path: str) -> None:  # myyuv_bmp.cpp:170-181
        try:
            with open(path, "wb") as f:
                f.write(self.header.pack())
                if self.header.bit_count == 32:
                    f.write(self.color_header.pack())
                f.write(self.data[: self.imageSize()].tobytes())
        except OSError:
            raise RuntimeError("Error opening file to write " + path)


class YUV:
    """myyuv::YUV (myyuv_yuv.hpp:37-350) restricted to what the hot path touches."""

    class FourccFormats:
        UNKNOWN = 0
        IYUV = 0x56555949

    class Compressions:
        NONE = 0
        DCT = 1

    # plugin registries, filled below (myyuv_yuv.cpp:88,130,146)
    bmp_to_yuv_map: dict = {}
    compress_map: dict = {}
    decompress_map: dict = {}
    yuv_resolution_fraction_map = {0x56555949: (2, 2)}

    def __init__(self, src=None, fmt: int | None = None):
        self.header = YUVHeader()
        self.compression_params: np.ndarray | None = None
        self.data: np.ndarray | None = None
        if isinstance(src, BMP):
            self.load_bmp(src, fmt)
        elif src is not None:
            self.load(src)

    # -- accessors --
    def getFourccFormat(self) -> int:
        return self.header.fourcc_format

    def getCompression(self) -> int:
        return self.header.compression

    def getWidth(self) -> int:
        return self.header.width

    def getHeight(self) -> int:
        return self.header.height

    def getDataSize(self) -> int:
        return self.header.data_size

    def isCompressed(self) -> bool:
        return self.getCompression() != YUV.Compressions.NONE

    def getImageSize(self) -> int:  # myyuv_yuv.cpp:374-381 for 4:2:0
        if not YUV.isImplementedFormat(self.getFourccFormat()):
            raise RuntimeError("Error. Unimplemented format.")
        wh = (self.header.width * self.header.height) & 0xFFFFFFFF
        return (wh + 2 * (wh * 2 // 8)) & 0xFFFFFFFF

    @staticmethod
    def isImplementedFormat(fmt: int, compression: int = 0) -> bool:  # myyuv_yuv.cpp:264-276
        if fmt not in YUV.bmp_to_yuv_map or fmt not in YUV.yuv_resolution_fraction_map:
            return False
        if compression != YUV.Compressions.NONE:
            return fmt in YUV.compress_map.get(compression, {}) and fmt in YUV.decompress_map.get(compression, {})
        return True

    def isValidHeader(self) -> bool:  # myyuv_yuv.cpp:256-262
        h = self.header
        return (h.type == b"YU" and YUV.isImplementedFormat(h.fourcc_format, h.compression) and h.width > 0 and h.height > 0
                and h.data_pos >= 64 + h.compression_params_size and h.data_size > 0)

    def isValid(self) -> bool:  # myyuv_yuv.cpp:248-254
        h = self.header
        cp = self.compression_params
        return (self.data is not None and ((h.compression_params_size > 0 and cp is not None)
                                           or (h.compression == 0 and cp is None)
                                           or (h.compression_params_size == 0 and cp is None)) and self.isValidHeader())

    def getYUVPlanes(self):
        w, h = self.header.width, self.header.height
        return self.data[: w * h], self.data[w * h: w * h * 5 // 4], self.data[w * h * 5 // 4: w * h * 3 // 2]

    # -- operations --
    def compress(self, compression: int, params, params_size: int | None = None) -> "YUV":  # myyuv_yuv.cpp:454-467
        if self.getCompression() != YUV.Compressions.NONE:
            raise RuntimeError("Error already compressed")
        if compression not in YUV.compress_map:
            raise RuntimeError("Error this compression is unimplemented")
        comp = YUV.compress_map[compression]
        if self.getFourccFormat() not in comp:
            raise RuntimeError("Error compression for this format is unimplemented")
        params = np.asarray(params, np.uint8).reshape(-1)
        return comp[self.getFourccFormat()](self, params, params.size if params_size is None else params_size)

    def decompress(self) -> "YUV":  # myyuv_yuv.cpp:469-483
        compression = self.getCompression()
        if compression == YUV.Compressions.NONE:
            return self.copy()
        if compression not in YUV.decompress_map:
            raise RuntimeError("Error this decompression is unimplemented")
        comp = YUV.decompress_map[compression]
        if self.getFourccFormat() not in comp:
            raise RuntimeError("Error decompression for this format is unimplemented")
        return comp[self.getFourccFormat()](self)

    def copy(self) -> "YUV":
        r = YUV()
        r.header = self.header.copy()
        r.data = None if self.data is None else self.data.copy()
        r.compression_params = None if self.compression_params is None else self.compression_params.copy()
        return r

    def load(self, path: str) -> None:  # myyuv_yuv.cpp:485-510
        try:
            raw = open(path, "rb").read()
        except OSError:
            raise RuntimeError("Error opening file to read " + path)
        res = YUV()
        res.header = YUVHeader.unpack(raw)
        if not res.isValidHeader():
            raise RuntimeError("Error bad header " + path)
        h = res.header
        if h.compression_params_size > 0:
            res.compression_params = np.frombuffer(raw, np.uint8, h.compression_params_size, h.compression_params_pos).copy()
        pos = h.data_pos
        h.compression_params_pos = 64
        h.data_pos = 64 + h.compression_params_size
        if res.getCompression() == YUV.Compressions.NONE:
            h.data_size = res.getImageSize()
        res.data = np.frombuffer(raw[pos:pos + h.data_size].ljust(h.data_size, b"\0"), np.uint8).copy()
        self.header, self.compression_params, self.data = res.header, res.compression_params, res.data

    def load_bmp(self, bmp: BMP, fmt: int) -> None:  # myyuv_yuv.cpp:512-523
        if not bmp.isValid():
            raise RuntimeError("BMP is invalid")
        if fmt not in YUV.bmp_to_yuv_map:
            raise RuntimeError("Incorrect format")
        tmp = YUV.bmp_to_yuv_map[fmt](bmp)
        self.header, self.compression_params, self.data = tmp.header, tmp.compression_params, tmp.data

    def dump(self, path: str) -> None:  # myyuv_yuv.cpp:525-536
        try:
            with open(path, "wb") as f:
                f.write(self.header.pack())
                if self.compression_params is not None:
                    f.write(self.compression_params[: self.header.compression_params_size].tobytes())
                f.write(self.data[: self.header.data_size].tobytes())
        except OSError:
            raise RuntimeError("Error opening file to write " + path)


# ---- registry entries: the three slots the B200 path plugs into ----
def _bmp_to_iyuv(bmp: BMP) -> YUV:  # replaces myyuv_yuv.cpp:89-127
    # the reference asserts 32 bpp (:92, "TODO: test 24"); its Release build (NDEBUG) converts 24-bit files as B,G,R triplets
    if bmp.header.bit_count not in (24, 32):
        raise RuntimeError("Error. only 24-bit and 32-bit BMP are supported")
    w, h = bmp.trueWidth(), bmp.trueHeight()
    pixels = bmp.data
    if bmp.header.width > 0 and bmp.header.height > 0:
        bottom_up = True
    elif bmp.header.width > 0 and bmp.header.height < 0:
        bottom_up = False
    elif bmp.header.width < 0 and bmp.header.height > 0:
        # colorData() reverses the whole pixel sequence for this orientation (myyuv_bmp.cpp:89-94); rare, so the
        # reordering is done on the host like the C++ drop-in does (yuv_host.cpp) and the rows are then top-down
        pb = bmp.header.bit_count // 8
        pixels = np.ascontiguousarray(np.asarray(bmp.data, np.uint8)[: w * h * pb].reshape(-1, pb)[::-1]).reshape(-1)
        bottom_up = False
    else:
        raise RuntimeError("Unaccounted width and height sign")  # myyuv_bmp.cpp:100
    res = YUV()
    res.header.fourcc_format = YUV.FourccFormats.IYUV
    res.header.width, res.header.height = w, h
    res.header.data_size = (w * h * 3 // 2) & 0xFFFFFFFF
    res.header.data_pos = 64
    try:
        ctx = capi.default_context()
        res.data = (ctx.xrgb_to_iyuv if bmp.header.bit_count == 32 else ctx.bgr24_to_iyuv)(pixels, w, h, bottom_up)
    except capi.MyyuvError as e:
        raise RuntimeError(str(e)) from e
    return res


def _compress_dct_iyuv(yuv: YUV, params: np.ndarray, params_size: int) -> YUV:  # replaces myyuv_yuv.cpp:132-142 + DCT.cpp:371-430
    if params_size != 3:
        raise RuntimeError("Error compression: incorrect parameters count. 3 parameters required")
    res = YUV()
    res.header = yuv.header.copy()
    try:
        res.data = capi.default_context().compress(yuv.data, yuv.header.width, yuv.header.height, params[:3])
    except capi.MyyuvError as e:
        raise RuntimeError(str(e)) from e
    res.header.compression = YUV.Compressions.DCT
    res.header.compression_params_size = 3
    res.header.compression_params_pos = 64
    res.header.data_pos = 67
    res.header.data_size = int(res.data.size)
    res.compression_params = np.array(params[:3], np.uint8)
    return res


def _decompress_dct_iyuv(yuv: YUV) -> YUV:  # replaces myyuv_yuv.cpp:148-158 + DCT.cpp:432-488
    if yuv.header.compression_params_size != 3:
        raise RuntimeError("Error decompression: incorrect parameters count. 3 parameters required")
    res = YUV()
    res.header = yuv.header.copy()
    res.header.compression = YUV.Compressions.NONE
    res.header.compression_params_size = 0
    res.header.compression_params_pos = 0
    res.header.data_pos = 64
    res.header.data_size = yuv.getImageSize()
    try:
        res.data = capi.default_context().decompress(yuv.data[: yuv.header.data_size], yuv.header.width, yuv.header.height,
                                                     yuv.compression_params[:3])
    except capi.MyyuvError as e:
        raise RuntimeError(str(e)) from e
    return res


YUV.bmp_to_yuv_map[YUV.FourccFormats.IYUV] = _bmp_to_iyuv
YUV.compress_map[YUV.Compressions.DCT] = {YUV.FourccFormats.IYUV: _compress_dct_iyuv}
YUV.decompress_map[YUV.Compressions.DCT] = {YUV.FourccFormats.IYUV: _decompress_dct_iyuv}
