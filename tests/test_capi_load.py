"""No GPU needed: the CUDA library loads, exports every symbol include/myyuvb200.h declares, answers the pure
host queries, and refuses to work (loudly, no CPU fallback) when there is no device."""
import ctypes as C
import pathlib
import re
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "myyuvb200.h").read_text()
    return sorted(set(re.findall(r"MYYUVB_API\s+[\w\s\*]+?\b(myyuvb_\w+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    names = declared_symbols()
    assert len(names) >= 18
    assert sorted(pkg.capi.EXPORTS) == names


def test_library_exports_every_declared_symbol(pkg):
    lib = C.CDLL(str(pkg.library_path()))
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg.library_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (myyuvb_\w+)", out))
    assert exported == set(declared_symbols())  # and nothing else leaks from the C ABI namespace


def test_library_is_sm100a_only(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", str(pkg.library_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_fma_contraction_in_sass(pkg):
    """Bit-exactness needs every product and sum of the DCT and the colour conversion rounded separately.
    The packed kernels use FMUL2 for products and FFMA2(p, 1.0, acc) for sums (448 per 8x8 transform when fully
    unrolled), plus the true FMAs of the exact-division step in the encoder; a scalar FFMA anywhere would be a contraction."""
    sass = subprocess.run(["cuobjdump", "-sass", str(pkg.library_path())], capture_output=True, text=True).stdout

    def counts(kernel):
        body = sass.split(kernel)[1].split("Function :")[0]
        return {m: len(re.findall(r"\b" + m + r"\b", body)) for m in ("FMUL2", "FFMA2", "FFMA", "DFMA")}

    enc, dec, col = counts("dct_compress_kernel"), counts("dct_decompress_kernel"), counts("xrgb_to_iyuv_kernel")
    assert enc["FFMA"] == 0 and dec["FFMA"] == 0 and col["FFMA"] == 0
    # decoder: the full transform plus the triangular variants K = 4 and 7 (4 * sum_c (K - c - 1) + 32 * (K - 1) sums each)
    tri = sum(4 * (K * (K - 1) // 2) + 32 * (K - 1) for K in (4, 7))
    # encoder: 224 sums of the first product (unrolled) + 56 of the second, whose loop over the output column PAIR is rolled
    # (8 rows x 7 sums per iteration), + 16 FMAs of the exact-division step in that loop body (two per coefficient pair)
    assert enc["FFMA2"] == 224 + 56 + 16 and dec["FFMA2"] == 448 + tri
    assert 248 + 64 + 8 <= enc["FMUL2"] <= 256 + 64 + 8 and 700 <= dec["FMUL2"] <= 512 + sum(2 * K * (K + 1) + 32 * K for K in (4, 7))  # identical products may be shared (exact)


def test_device_code_identity(pkg):
    """bench.py reports the DRAM traffic of a committed ncu capture only for the machine code the capture was taken from:
    the identity is a hash of the library's SASS, which a comment edit keeps and a changed instruction does not."""
    import importlib, json

    build = importlib.import_module("yuv-manipulations-2_b200.build")
    a, b = build.device_code_sha256(), build.device_code_sha256(pkg.library_path())
    assert a is not None and re.fullmatch(r"[0-9a-f]{64}", a) and a == b
    assert build.device_code_sha256(pathlib.Path(__file__)) is None  # not a CUDA binary: no identity, bench falls back to the source hash
    tr = json.loads((pathlib.Path(__file__).resolve().parent.parent / "profiles" / "r02_traffic.json").read_text())
    assert re.fullmatch(r"[0-9a-f]{64}", tr["device_code_sha256"]) and re.fullmatch(r"[0-9a-f]{64}", tr["sources_sha256"])


def test_device_free_entry_points(pkg):
    """The C ABI's entry points that need no device: the band split of a sharded image (270 macroblock rows of 8K over
    8 ranks = 34 x 6 + 33 x 2) agrees with the host-side split the gloo tests use, plane accessors follow YUV::getYUVPlanes /
    getWidthHeightChannel (myyuv_yuv.cpp:383-427), and arguments are rejected with the documented codes before anything is launched."""
    import importlib

    capi = pkg.capi
    sh = importlib.import_module("yuv-manipulations-2_b200.sharding")
    rows = capi.shard_rows(4320, 8)
    assert [(b - a) // 16 for a, b in zip(rows, rows[1:])] == [34, 34, 34, 34, 34, 34, 33, 33] and rows[0] == 0 and rows[-1] == 4320
    for h in (16, 32, 48, 736, 2160 // 16 * 16, 4320):
        for world in (1, 2, 3, 5, 8, 16):
            rows = capi.shard_rows(h, world)
            assert list(zip(rows, rows[1:])) == sh.macroblock_row_bands(h, world)
    for bad in ((4328, 8), (4320, 0), (4320, 17)):
        with pytest.raises(pkg.MyyuvError) as e:
            capi.shard_rows(*bad)
        assert e.value.code == (capi.ERR_HEIGHT if bad[0] % 16 else capi.ERR_ARG)
    base = 0x10000
    assert capi.iyuv_planes(base, 3840, 2160) == [(base, 3840, 2160), (base + 3840 * 2160, 1920, 1080), (base + 3840 * 2160 * 5 // 4, 1920, 1080)]
    assert capi.shard_ctrl_bytes() >= 16 * 16 + 16 * 4 and capi.shard_ctrl_bytes() % 8 == 0
    assert capi.compress_bound(3840, 2160) == 12 + 3 * 8 + (3840 * 2160 // 64 * 3 // 2) * 256  # headers + 1 size byte + 255 chunk bytes per block


def test_compress_bound(pkg):
    nblk = (3840 // 8) * (2160 // 8) * 3 // 2
    assert pkg.capi.compress_bound(3840, 2160) == 36 + nblk * 256


def test_no_device_no_fallback(pkg):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.MyyuvError, match="no usable CUDA device"):
        pkg.Context(0)
    # the class API goes through the same path and must fail as well, not compute on the CPU
    y = pkg.YUV()
    y.header.fourcc_format = pkg.YUV.FourccFormats.IYUV
    y.header.width = y.header.height = 16
    y.header.data_size = 384
    y.header.data_pos = 64
    y.data = np.zeros(384, np.uint8)
    with pytest.raises((RuntimeError, pkg.MyyuvError)):
        y.compress(pkg.YUV.Compressions.DCT, [50, 50, 50])


def test_class_api_host_logic(pkg, tmp_path):
    """Header packing, load/dump and the checks that need no device."""
    YUV = pkg.YUV
    y = YUV()
    y.header.fourcc_format = YUV.FourccFormats.IYUV
    y.header.width, y.header.height, y.header.data_size, y.header.data_pos = 16, 16, 384, 64
    y.data = np.arange(384, dtype=np.uint8)
    assert y.isValid() and not y.isCompressed() and y.getImageSize() == 384
    p = tmp_path / "a.myyuv"
    y.dump(str(p))
    assert p.stat().st_size == 64 + 384
    z = YUV(str(p))
    assert z.isValid() and np.array_equal(z.data, y.data)
    # load() normalises the positions like the reference (myyuv_yuv.cpp:501-502)
    assert (z.header.width, z.header.height, z.header.data_size, z.header.compression_params_pos, z.header.data_pos) == (16, 16, 384, 64, 64)
    assert np.array_equal(z.decompress().data, y.data)  # uncompressed -> copy (myyuv_yuv.cpp:471-473)
    c = z.copy()
    c.header.compression = YUV.Compressions.DCT
    with pytest.raises(RuntimeError, match="Error already compressed"):
        c.compress(YUV.Compressions.DCT, [50, 50, 50])
    with pytest.raises(RuntimeError, match="Error this compression is unimplemented"):
        z.compress(7, [50, 50, 50])
    with pytest.raises(RuntimeError, match="Error opening file to read"):
        YUV(str(tmp_path / "missing.myyuv"))
    (tmp_path / "bad.myyuv").write_bytes(b"XX" + bytes(100))
    with pytest.raises(RuntimeError, match="Error bad header"):
        YUV(str(tmp_path / "bad.myyuv"))
