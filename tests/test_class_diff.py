"""Host-side behaviour of the drop-in class library against the UNMODIFIED reference library, without a GPU: the same C++ program
(tests/hostemu/class_diff.cpp) linked against each prints what myyuv::BMP / myyuv::YUV do with a set of well-formed and malformed
files -- header normalisation on load (myyuv_bmp.cpp:141-166, myyuv_yuv.cpp:485-510), validity rules (:127-139, :248-262),
orientation handling of colorData() (:80-103), accessors, getPixel (myyuv_yuv.cpp:162-180), copies / moves, dump, exception texts --
and the two transcripts must be identical.  What reaches the codec is covered by the GPU tests (tests/test_dropin_cli.py)."""
import os
import pathlib
import struct
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
HERE = ROOT / "tests" / "hostemu"
REF_LIB = ROOT / "oracle" / "_ref" / "serial" / "libmyyuv_lib.so"
GOLD = ROOT / "oracle" / "_ref" / "golden"

BMP_HDR = struct.Struct("<2sIHHIIiiHHIIiiII")   # myyuv_bmp.hpp:12-31, 54 bytes
BMP_COL = struct.Struct("<IIIII64s")            # myyuv_bmp.hpp:33-43, 84 bytes
YUV_HDR = struct.Struct("<2sIIHIIIII32s")       # myyuv_yuv.hpp:13-29, 64 bytes


def bmp_file(w, h, bits=32, gap=0, typ=b"BM", compression=None, used=0, important=0, masks=(0x00FF0000, 0x0000FF00, 0x000000FF, 0xFF000000),
             space=0x73524742, header_size=None, seed=0):
    rng = np.random.default_rng(seed)
    pix = rng.integers(0, 256, abs(w) * abs(h) * bits // 8, dtype=np.uint8).tobytes()
    has_masks = bits == 32
    data_pos = 54 + (84 if has_masks else 0) + gap
    hdr = BMP_HDR.pack(typ, data_pos + len(pix), 0, 0, data_pos, (124 if has_masks else 40) if header_size is None else header_size, w, h, 1, bits,
                       (3 if has_masks else 0) if compression is None else compression, len(pix), 2835, 2835, used, important)
    col = BMP_COL.pack(*masks, space, bytes(64)) if has_masks else b""
    return hdr + col + bytes(range(gap)) + pix


def yuv_file(w, h, gap=0, typ=b"YU", fourcc=0x56555949, compression=0, params=b"", data=None, seed=0):
    rng = np.random.default_rng(seed)
    if data is None:
        data = rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8).tobytes()
    ppos = 64 + gap if params else 0
    dpos = 64 + gap + len(params) + gap
    hdr = YUV_HDR.pack(typ, fourcc, len(data), compression, len(params), ppos, w, h, dpos, bytes(32))
    return hdr + bytes(gap) + params + bytes(gap) + data


@pytest.fixture(scope="module")
def binaries():
    if not REF_LIB.exists():
        pytest.skip("oracle/_ref (the unmodified reference library) is not built")
    subprocess.run(["make", "-s", "-C", str(HERE), "class_diff"], check=True)
    return HERE / "class_diff_ours", HERE / "class_diff_ref"


def write_cases(tmp_path):
    files = {
        "a_32_bottom_up.bmp": bmp_file(8, 6, seed=1),
        "b_32_top_down.bmp": bmp_file(8, -6, seed=2),
        "c_32_mirrored.bmp": bmp_file(-8, 6, seed=3),
        "d_32_both_negative.bmp": bmp_file(-8, -6, seed=4),
        "e_32_gap_before_pixels.bmp": bmp_file(12, 4, gap=10, seed=5),
        "f_24_bottom_up.bmp": bmp_file(8, 6, bits=24, seed=6),
        "g_24_top_down.bmp": bmp_file(4, -2, bits=24, seed=7),
        "h_24_mirrored.bmp": bmp_file(-4, 2, bits=24, seed=8),
        "i_alpha_mask_zero.bmp": bmp_file(4, 4, masks=(0x00FF0000, 0x0000FF00, 0x000000FF, 0), seed=9),
        "j_bi_rgb_32.bmp": bmp_file(4, 4, compression=0, seed=10),
        "k_bad_type.bmp": bmp_file(4, 4, typ=b"XM"),
        "l_bad_width.bmp": bmp_file(6, 4),
        "m_bad_compression.bmp": bmp_file(4, 4, compression=1),
        "n_bad_colors_used.bmp": bmp_file(4, 4, used=2),
        "o_bad_colors_important.bmp": bmp_file(4, 4, important=1),
        "p_bad_masks.bmp": bmp_file(4, 4, masks=(0x000000FF, 0x0000FF00, 0x00FF0000, 0xFF000000)),
        "q_bad_alpha.bmp": bmp_file(4, 4, masks=(0x00FF0000, 0x0000FF00, 0x000000FF, 0x0F000000)),
        "r_bad_space.bmp": bmp_file(4, 4, space=0x57696E20),
        "s_bad_header_size.bmp": bmp_file(4, 4, header_size=0),
        "t_16_bit.bmp": bmp_file(4, 4, bits=16),
        "u_iyuv.myyuv": yuv_file(16, 16, seed=11),
        "v_iyuv_gap.myyuv": yuv_file(32, 16, gap=16, seed=12),
        "w_compressed_header_only.myyuv": yuv_file(16, 16, compression=1, params=bytes([50, 60, 70]), data=bytes(range(200)), seed=13),
        "x_bad_type.myyuv": yuv_file(16, 16, typ=b"YX"),
        "y_unknown_fourcc.myyuv": yuv_file(16, 16, fourcc=0x32595559),
        "z_unknown_compression.myyuv": yuv_file(16, 16, compression=9, params=b"\x01", data=bytes(64)),
    }
    paths = []
    for name, blob in files.items():
        (tmp_path / name).write_bytes(blob)
        paths.append(str(tmp_path / name))
    paths.append(str(tmp_path / "missing.bmp"))
    paths.append(str(tmp_path / "missing.myyuv"))
    for g in ("chef-with-trumpet.bmp", "chef-with-trumpet.myyuv", "chef-with-trumpet-DCT-50.myyuv"):
        if (GOLD / g).exists():
            paths.append(str(GOLD / g))
    return paths


def test_class_api_host_behaviour_equals_the_reference(binaries, tmp_path):
    paths = write_cases(tmp_path)
    outs = []
    for i, exe in enumerate(binaries):
        scratch = tmp_path / f"scratch{i}"
        scratch.mkdir()
        r = subprocess.run([str(exe), str(scratch)] + paths, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.splitlines())
    ours, ref = outs
    assert len(ref) > 150  # the transcript is not trivially empty
    diff = [(a, b) for a, b in zip(ours, ref) if a != b]
    assert len(ours) == len(ref) and not diff, "first differences (ours, reference):\n" + "\n".join(f"{a}\n{b}" for a, b in diff[:8])


def test_class_api_host_side_under_sanitizers(tmp_path):
    """The same transcript with the drop-in library's host classes compiled under AddressSanitizer + UBSan, plus the one query the
    differential run leaves out because the reference reads past its buffer there (getPixel on the last row, right half)."""
    r = subprocess.run(["make", "-s", "-C", str(HERE), str(HERE / "class_diff_asan")], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("no sanitizer runtime for /usr/bin/g++")
    assert r.returncode == 0, r.stderr[-2000:]
    paths = write_cases(tmp_path)
    scratch = tmp_path / "scratch"
    scratch.mkdir()
    env = dict(os.environ, CLASS_DIFF_EXTRA="1", ASAN_OPTIONS="detect_leaks=1:protect_shadow_gap=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([str(HERE / "class_diff_asan"), str(scratch)] + paths, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    assert "pixel (last row, right half)" in r.stdout and len(r.stdout.splitlines()) > 150
