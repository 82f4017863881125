"""Pins the CPU oracle (oracle/myyuv_oracle.c): against the committed golden hashes that were produced by the
UNMODIFIED reference (tests/golden/make_golden.py), against the reference's own sample images when they are
staged (oracle/_ref/golden), and against the reference library itself when it was built here (oracle/_ref)."""
import hashlib
import json
import pathlib

import numpy as np
import pytest

GOLDEN = json.loads((pathlib.Path(__file__).parent / "golden" / "golden.json").read_text())
GDIR = pathlib.Path(__file__).parent / "golden"


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", [c for c in GOLDEN["synthetic"] if c["w"] <= 1920], ids=lambda c: f"{c['w']}x{c['h']}q{c['q']}")
def test_compress_decompress_match_reference_hashes(ora, synth, case):
    w, h, q = case["w"], case["h"], case["q"]
    f = synth.iyuv_frames_numpy(w, h, 1, case["first"])[0]
    assert sha(f) == case["input_sha256"], "synthetic generator changed"
    c = ora.compress(f, w, h, q)
    assert c.size == case["payload_size"] and sha(c) == case["payload_sha256"]
    assert sha(ora.decompress(c, w, h, q)) == case["decoded_sha256"]


@pytest.mark.parametrize("case", GOLDEN["colour"], ids=lambda c: f"{c['w']}x{c['h']}bu{int(c['bottom_up'])}")
def test_colour_matches_reference_hashes(ora, synth, case):
    b = synth.bgrx_frames_numpy(case["w"], case["h"], 1, case["first"])[0]
    assert sha(b) == case["input_sha256"]
    assert sha(ora.bgrx_to_iyuv(b, case["w"], case["h"], case["bottom_up"])) == case["iyuv_sha256"]


def bgr24_frame(synth, w, h, first):
    return np.ascontiguousarray(synth.bgrx_frames_numpy(w, h, 1, first)[0].reshape(-1, 4)[:, :3]).reshape(-1)


@pytest.mark.parametrize("case", GOLDEN["colour24"], ids=lambda c: f"{c['w']}x{c['h']}bu{int(c['bottom_up'])}")
def test_colour24_matches_reference_hashes(ora, synth, case):
    """24-bit BMP rows: what the reference's Release build (assert compiled out) makes of B,G,R triplets."""
    b = bgr24_frame(synth, case["w"], case["h"], case["first"])
    assert sha(b) == case["input_sha256"]
    got = ora.bgr24_to_iyuv(b, case["w"], case["h"], case["bottom_up"])
    assert sha(got) == case["iyuv_sha256"]
    # the X byte never enters the arithmetic: same planes as the 32-bit frame the triplets were cut from
    full = synth.bgrx_frames_numpy(case["w"], case["h"], 1, case["first"])[0]
    assert np.array_equal(got, ora.bgrx_to_iyuv(full, case["w"], case["h"], case["bottom_up"]))


def test_colour24_oracle_against_reference_itself(ora, ref):
    rng = np.random.default_rng(24)
    for w, h in ((4, 2), (12, 6), (40, 18)):
        b = rng.integers(0, 256, w * h * 3, dtype=np.uint8)
        for bottom_up in (True, False):
            assert np.array_equal(ora.bgr24_to_iyuv(b, w, h, bottom_up), ref.bgr24_to_iyuv(b, w, h, bottom_up))


@pytest.mark.parametrize("case", GOLDEN["edge"], ids=lambda c: f"q{c['q']}")
def test_edge_cases_match_reference_hashes(ora, synth, case):
    f = synth.edge_case_iyuv(case["w"], case["h"])
    assert sha(f) == case["input_sha256"]
    c = ora.compress(f, case["w"], case["h"], case["q"])
    assert sha(c) == case["payload_sha256"]
    assert sha(ora.decompress(c, case["w"], case["h"], case["q"])) == case["decoded_sha256"]


def test_tiny_fixture_bytes(ora):
    f = np.load(GDIR / "tiny_32x32_q50_input.npy")
    c = np.load(GDIR / "tiny_32x32_q50_payload.npy")
    d = np.load(GDIR / "tiny_32x32_q50_decoded.npy")
    assert np.array_equal(ora.compress(f, 32, 32, (50, 50, 50)), c)
    assert np.array_equal(ora.decompress(c, 32, 32, (50, 50, 50)), d)


def golden_blocks():
    g = GOLDEN["huffman_blocks"]
    rng = np.random.default_rng(g["seed"])
    blocks = np.zeros((g["n"], 64), np.int16)
    for i in range(g["n"]):
        m = int(rng.integers(1, 65))
        vals = rng.choice(np.arange(-1024, 1024), m, replace=False)
        blocks[i] = vals[rng.integers(0, m, 64)]
        if i % 3 == 0:
            blocks[i, rng.integers(0, 64, 40)] = 0
    return blocks, g


def test_huffman_tie_breaking_matches_reference_hashes(ora):
    """4096 random blocks with 1..64 distinct symbols: exercises every rehash step of the libstdc++ emulation."""
    blocks, g = golden_blocks()
    chunks, sizes = ora.huff_encode_blocks(blocks)
    assert sha(sizes) == g["sizes_sha256"] and sha(chunks) == g["chunks_sha256"]
    assert np.array_equal(ora.huff_decode_blocks(chunks, sizes), blocks)


def test_reference_sample_images(ora, golden_dir):
    import oracle as O

    for name, digest in GOLDEN["chef"].items():
        if not name.startswith("decoded:"):
            assert hashlib.sha256((golden_dir / name).read_bytes()).hexdigest() == digest
    bmp = O.read_bmp32(golden_dir / "chef-with-trumpet.bmp")
    raw = O.read_myyuv(golden_dir / "chef-with-trumpet.myyuv")
    assert np.array_equal(ora.bgrx_to_iyuv(bmp["data"], bmp["w"], bmp["h"], bmp["bottom_up"]), raw["data"])
    for q in (50, 90):
        g = O.read_myyuv(golden_dir / f"chef-with-trumpet-DCT-{q}.myyuv")
        assert list(g["params"]) == [q, q, q]
        assert np.array_equal(ora.compress(raw["data"], raw["w"], raw["h"], [q] * 3), g["data"])
        assert sha(ora.decompress(g["data"], g["w"], g["h"], g["params"])) == GOLDEN["chef"][f"decoded:chef-with-trumpet-DCT-{q}.myyuv"]


def test_reference_big_image_decode(ora, golden_dir):
    import oracle as O

    g = O.read_myyuv(golden_dir / "chef-with-trumpet-big-DCT-50.myyuv")
    assert (g["w"], g["h"]) == (4032, 3008)
    assert sha(ora.decompress(g["data"], g["w"], g["h"], g["params"])) == GOLDEN["chef"]["decoded:chef-with-trumpet-big-DCT-50.myyuv"]


def test_against_reference_library(ora, ref, synth):
    rng = np.random.default_rng(3)
    for w, h, q in [(64, 64, (33, 66, 99)), (320, 240 - 240 % 16, (5, 5, 5)), (256, 256, (100, 100, 100))]:
        f = rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8) if q[0] == 100 else synth.iyuv_frames_numpy(w, h, 1, 5)[0]
        c = ora.compress(f, w, h, q)
        assert np.array_equal(c, ref.compress(f, w, h, q))
        assert np.array_equal(ora.decompress(c, w, h, q), ref.decompress(c, w, h, q))
    blocks = rng.integers(-1024, 1024, (2000, 64)).astype(np.int16)
    blocks[:, 20:] = np.where(rng.random((2000, 44)) < 0.8, 0, blocks[:, 20:])
    co, so = ora.huff_encode_blocks(blocks)
    cr, sr = ref.huff_encode_blocks(blocks)
    assert np.array_equal(so, sr) and np.array_equal(co, cr)
    assert np.array_equal(ref.huff_decode_blocks(co, so), blocks)


def test_oracle_error_codes(ora, synth):
    import oracle as O

    f = synth.iyuv_frames_numpy(32, 32, 1)[0]
    with pytest.raises(O.OracleError) as e:
        ora.compress(f, 32, 32, (0, 50, 50))
    assert e.value.code == O.Oracle.ERR_QUALITY
    good = ora.compress(f, 32, 32, (50, 50, 50))
    with pytest.raises(O.OracleError) as e:
        ora.decompress(good[:12], 32, 32, (50, 50, 50))
    assert e.value.code == O.Oracle.ERR_DCTYUV_SIZE
    bad = good.copy()
    bad[12 + 8 + 16] = 0xFF
    bad[12 + 8 + 17] = 0x01
    with pytest.raises(O.OracleError):
        ora.decompress(bad, 32, 32, (50, 50, 50))


def test_qtable_formula(ora):
    lum50 = ora.qtable(50, False)
    assert lum50[0] == 16 and lum50[63] == 99
    assert np.all(ora.qtable(100, True) == 1)
    q1 = ora.qtable(1, False)
    assert q1.max() == 255 and q1.min() == 255  # 50/1 * 10 = 500 -> clamped
    assert ora.qtable(75, False)[0] == 8


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_natural_content_4k_matches_reference_hashes(ora, synth, golden_dir, idx):
    """tiled-real 4K frames at q 50 / 90 / 10 against the reference's hashes."""
    import oracle

    case = GOLDEN["tiled_real"][idx]
    g = oracle.read_myyuv(golden_dir / "chef-with-trumpet.myyuv")
    w, h, q = case["w"], case["h"], tuple(case["q"])
    f = synth.tiled_real_iyuv(g["data"], g["w"], g["h"], w, h, 1, case["first"])[0]
    assert sha(f) == case["input_sha256"]
    c = ora.compress(f, w, h, q)
    assert c.size == case["payload_size"] and sha(c) == case["payload_sha256"]
    assert sha(ora.decompress(c, w, h, q)) == case["decoded_sha256"]


def test_tiled_real_generator(synth):
    """natural frames for the side measurements are a base image tiled with a per-frame shift."""
    rng = np.random.default_rng(0)
    bw, bh, w, h = 32, 16, 80, 48
    base = rng.integers(0, 256, bw * bh * 3 // 2, dtype=np.uint8)
    fr = synth.tiled_real_iyuv(base, bw, bh, w, h, 3, first=1)
    assert fr.shape == (3, w * h * 3 // 2)
    Y = base[: bw * bh].reshape(bh, bw)
    y1 = fr[0, : w * h].reshape(h, w)
    assert y1[0, 0] == Y[16 % bh, 16 % bw] and y1[5, 7] == Y[(5 + 16) % bh, (7 + 16) % bw]
    U = base[bw * bh: bw * bh * 5 // 4].reshape(bh // 2, bw // 2)
    u1 = fr[0, w * h: w * h * 5 // 4].reshape(h // 2, w // 2)
    assert u1[3, 2] == U[(3 + 8) % (bh // 2), (2 + 8) % (bw // 2)]
    assert np.array_equal(fr, synth.tiled_real_iyuv(base, bw, bh, w, h, 3, first=1))
