"""The product's per-block entropy coder (csrc/block_codec.cuh, the source the CUDA kernels inline) compiled for
the host and compared with the oracle -- runs without a GPU.  Also checks the two arithmetic identities the
kernels rely on for bit-exactness: exact division by a single Newton step, rounding via a round-toward-zero add."""
import ctypes as C
import pathlib
import subprocess

import numpy as np
import pytest

HERE = pathlib.Path(__file__).parent / "hostemu"
u8p, i16p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16)


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    import os

    # HOSTEMU_LIB: another build of the same source, e.g. the AddressSanitizer + UBSan build test_hostemu_under_sanitizers runs this file with
    L = C.CDLL(os.environ.get("HOSTEMU_LIB") or str(HERE / "libhostemu.so"))
    L.hostemu_encode_blocks.argtypes = [i16p, C.c_uint32, C.c_int, C.c_int, u8p, u8p]
    L.hostemu_decode_blocks.argtypes = [u8p, u8p, C.c_uint32, i16p]
    L.hostemu_decode_blocks2.argtypes = [u8p, u8p, C.c_uint32, C.c_int, i16p, C.POINTER(C.c_uint32)]
    L.hostemu_division_check.argtypes = [C.c_uint32, C.c_uint32]
    L.hostemu_division_check.restype = C.c_uint64
    L.hostemu_round_check.argtypes = [C.c_uint32, C.c_uint32]
    L.hostemu_round_check.restype = C.c_uint64
    return L


def make_blocks(kind, n, rng):
    b = np.zeros((n, 64), np.int16)
    if kind == "sparse":
        for i in range(n):
            k = rng.integers(0, 20)
            b[i, rng.integers(0, 64, k)] = rng.integers(-6, 7, k)
    elif kind == "mid":
        for i in range(n):
            k = rng.integers(0, 40)
            b[i, rng.integers(0, 64, k)] = rng.integers(-40, 41, k)
    elif kind == "distinct":
        for i in range(n):
            m = rng.integers(1, 65)
            vals = rng.choice(np.arange(-1024, 1024), m, replace=False)
            b[i] = vals[rng.integers(0, m, 64)]
    elif kind == "full":
        b = rng.integers(-1024, 1024, (n, 64)).astype(np.int16)
    elif kind == "small":
        b = rng.integers(-3, 4, (n, 64)).astype(np.int16)
    elif kind == "nozero":  # no zero in the message and exactly 13/29/59 distinct symbols: the appended-key-0 rehash
        nz = np.concatenate([np.arange(-1024, 0), np.arange(1, 1024)])
        for i in range(n):
            m = [13, 29, 59, 12, 14, 28, 30, 58, 60][i % 9]
            vals = rng.choice(nz, m, replace=False)
            seq = np.concatenate([vals, vals[rng.integers(0, m, 64 - m)]])
            rng.shuffle(seq)
            b[i] = seq
    elif kind == "few":  # 1..17 distinct symbols (the fast path, its rehash replay and its boundary), with frequency ties and hash collisions
        for i in range(n):
            m = rng.integers(1, 18)
            pool = [np.arange(-4, 5), np.arange(-1024, 1024), np.arange(-60, 61), np.arange(-13 * 6, 13 * 6 + 1, 13),
                    np.arange(-32 * 8, 32 * 8 + 1, 32)][i % 5]
            vals = rng.choice(pool, min(m, pool.size), replace=False)
            L = rng.integers(1, 65)
            b[i, :L] = vals[rng.integers(0, vals.size, L)]
    elif kind == "around32":  # the boundary between heavy_blocks_kernel's 32-symbol and 64-symbol scratch
        for i in range(n):
            m = 30 + i % 6
            pool = np.arange(-70, 71) if i % 2 else np.arange(-1024, 1024)
            vals = rng.choice(pool, m, replace=False)
            seq = np.concatenate([vals, vals[rng.integers(0, m, 64 - m)]])
            rng.shuffle(seq)
            b[i] = seq
            if i % 3 == 0:
                b[i, 50 + i % 14:] = 0
    elif kind == "zeros":
        pass
    return b


@pytest.mark.parametrize("kind,n", [("sparse", 6000), ("mid", 6000), ("distinct", 6000), ("full", 1500), ("small", 6000),
                                    ("nozero", 1800), ("zeros", 64), ("few", 40000), ("around32", 3000)])
@pytest.mark.parametrize("stride,fast", [(1, 1), (128, 1), (1, 0), (1, 2), (32, 2), (1, 3), (64, 3), (32, 4), (1, 4), (32, 5)])
def test_block_coder_matches_oracle(emu, ora, kind, n, stride, fast):
    rng = np.random.default_rng(hash(kind) % 1000)
    b = make_blocks(kind, n, rng)
    want_c, want_s = ora.huff_encode_blocks(b)
    out = np.empty(n * 256, np.uint8)
    sizes = np.empty(n, np.uint8)
    emu.hostemu_encode_blocks(b.ctypes.data_as(i16p), n, stride, fast, out.ctypes.data_as(u8p), sizes.ctypes.data_as(u8p))
    assert np.array_equal(sizes, want_s)
    assert np.array_equal(out[: int(sizes.sum(dtype=np.int64))], want_c)
    dec = np.empty((n, 64), np.int16)
    rc = emu.hostemu_decode_blocks(want_c.ctypes.data_as(u8p), want_s.ctypes.data_as(u8p), n, dec.ctypes.data_as(i16p))
    assert rc == 0 and np.array_equal(dec, b)
    dec2 = np.empty((n, 64), np.int16)
    used = C.c_uint32(0)
    rc = emu.hostemu_decode_blocks2(want_c.ctypes.data_as(u8p), want_s.ctypes.data_as(u8p), n, 1, dec2.ctypes.data_as(i16p), C.byref(used))
    assert rc == 0 and np.array_equal(dec2, b)
    if kind in ("sparse", "small", "few", "zeros"):
        assert used.value > n // 2  # the fast decoder really ran


def test_decoder_rejects_truncated_stream(emu, ora):
    b = make_blocks("mid", 4, np.random.default_rng(1))
    c, s = ora.huff_encode_blocks(b)
    bad = c.copy()
    bad[0] = 0xFF
    bad[1] = 0x01  # 511 code bits in a chunk that holds far fewer
    dec = np.empty((4, 64), np.int16)
    assert emu.hostemu_decode_blocks(bad.ctypes.data_as(u8p), s.ctypes.data_as(u8p), 4, dec.ctypes.data_as(i16p)) == 1
    assert emu.hostemu_decode_blocks2(bad.ctypes.data_as(u8p), s.ctypes.data_as(u8p), 4, 1, dec.ctypes.data_as(i16p), None) == 1


def test_fast_and_general_decoder_agree_on_damaged_chunks(emu, ora):
    """Flip bytes of valid chunks: both decoders must report an error for the same blocks and produce the same
    coefficients for the blocks they accept (the fast one declines what it cannot reproduce)."""
    rng = np.random.default_rng(7)
    b = make_blocks("few", 3000, rng)
    c, s = ora.huff_encode_blocks(b)
    off = np.concatenate([[0], np.cumsum(s, dtype=np.int64)])
    bad = c.copy()
    for i in range(len(s)):
        k = rng.integers(0, s[i])
        bad[off[i] + k] ^= 1 << rng.integers(0, 8)
    for i in range(len(s)):
        one, sz = bad[off[i]: off[i + 1]].copy(), s[i: i + 1].copy()
        d0, d1 = np.zeros((1, 64), np.int16), np.zeros((1, 64), np.int16)
        r0 = emu.hostemu_decode_blocks2(one.ctypes.data_as(u8p), sz.ctypes.data_as(u8p), 1, 0, d0.ctypes.data_as(i16p), None)
        r1 = emu.hostemu_decode_blocks2(one.ctypes.data_as(u8p), sz.ctypes.data_as(u8p), 1, 1, d1.ctypes.data_as(i16p), None)
        assert r0 == r1, i
        if r0 == 0:
            assert np.array_equal(d0, d1), i


def test_decoders_agree_on_arbitrary_bytes(emu, ora):
    """Chunks that no encoder wrote: random bytes of every size from 0 to 255, random bytes behind a plausible header, and valid
    chunks cut short.  The step-by-step decoder (Huffman.cpp:106-154, :243-277) and the kernels' fast decoder must give the same
    verdict and, where they accept, the same coefficients; the oracle's decoder must agree with both (it is the reference's
    behaviour restated).  Each chunk sits in a buffer of exactly its size, so the sanitizer run sees any read past it."""
    rng = np.random.default_rng(11)
    valid_c, valid_s = ora.huff_encode_blocks(make_blocks("few", 600, rng))
    voff = np.concatenate([[0], np.cumsum(valid_s, dtype=np.int64)])
    cases = []
    for size in range(0, 256):
        for _ in range(8):
            cases.append(rng.integers(0, 256, size, dtype=np.uint8))
    for _ in range(4000):  # header says: few bits, small table -- the rest is noise
        size = int(rng.integers(3, 80))
        c = rng.integers(0, 256, size, dtype=np.uint8)
        c[0], c[1], c[2] = rng.integers(0, 256), rng.integers(0, 2), rng.integers(0, size)
        cases.append(c)
    for i in range(600):  # truncated valid chunks
        full = valid_c[voff[i]: voff[i + 1]]
        cases.append(full[: int(rng.integers(0, full.size))].copy())
    accepted = 0
    for i, c in enumerate(cases):
        one = np.ascontiguousarray(c)
        if one.size == 0:
            one = np.zeros(1, np.uint8)[:0].copy()
        sz = np.array([c.size], np.uint8)
        d0, d1 = np.zeros((1, 64), np.int16), np.zeros((1, 64), np.int16)
        buf = one if one.size else np.zeros(1, np.uint8)  # a pointer is needed even for an empty chunk
        r0 = emu.hostemu_decode_blocks2(buf.ctypes.data_as(u8p), sz.ctypes.data_as(u8p), 1, 0, d0.ctypes.data_as(i16p), None)
        r1 = emu.hostemu_decode_blocks2(buf.ctypes.data_as(u8p), sz.ctypes.data_as(u8p), 1, 1, d1.ctypes.data_as(i16p), None)
        assert r0 == r1, (i, c.size)
        try:
            want = ora.huff_decode_blocks(buf[: c.size], sz) if c.size else None
        except Exception:  # noqa: BLE001 -- the oracle raises where the reference throws
            want = None
        assert (want is not None) == (r0 == 0) or c.size == 0, (i, c.size)
        if r0 == 0:
            accepted += 1
            assert np.array_equal(d0, d1), i
            if want is not None:
                assert np.array_equal(want.reshape(1, 64), d0), i
    assert 0 < accepted < len(cases)  # the family contains both kinds


def test_exact_division_identity(emu):
    assert emu.hostemu_division_check(40000, 1) == 0


def test_round_half_away_identity(emu):
    assert emu.hostemu_round_check(400000, 2) == 0


def test_block_row_col_identity():
    """kernels.cu block_row_col: block index -> (block row, block column) without a division.  With magic = floor(2^32 / bw),
    floor(k * magic / 2^32) is the row or one less for every k < 2^32 (the error of the product is below k / 2^32 < 1), and one
    conditional step repairs it; bw = 1 uses magic = 2^32 - 1.  Checked on every plane width of the tested image sizes, on
    awkward widths, at the row boundaries and at the top of the 32-bit range."""
    rng = np.random.default_rng(5)
    for bw in (1, 2, 3, 5, 7, 62, 124, 240, 252, 480, 504, 960, 1023, 8191, 65535):
        magic = (1 << 32) // bw if bw > 1 else 0xFFFFFFFF
        q = rng.integers(0, (1 << 32) // bw, 20000, dtype=np.uint64)
        k = np.concatenate([q * bw, q * bw + (bw - 1), np.minimum(q * bw + rng.integers(0, bw, q.size, dtype=np.uint64), (1 << 32) - 1),
                            np.array([0, 1, bw - 1, bw, (1 << 32) - 1, (1 << 32) - bw], np.uint64)])
        k = k[k < (1 << 32)]
        by = (k * np.uint64(magic)) >> np.uint64(32)
        bx = k - by * np.uint64(bw)
        fix = bx >= bw
        by, bx = by + fix, bx - fix * np.uint64(bw)
        assert np.array_equal(by, k // np.uint64(bw)) and np.array_equal(bx, k % np.uint64(bw))


def test_hostemu_under_sanitizers():
    """Every test of this file once more against the AddressSanitizer + UBSan build of the same source: the coder's and the decoders'
    index arithmetic (scratch rows, nibble lists, heap, bit windows) on 70 000 blocks and 3 000 damaged chunks without an
    out-of-bounds access, an over-wide shift or a signed overflow.  (What the sanitizers cannot see -- an index that stays inside
    the scratch allocation but leaves its row -- is what the byte-exact comparison with the oracle catches.)"""
    import os
    import sys

    if os.environ.get("HOSTEMU_LIB"):
        pytest.skip("already the sanitizer run")
    r = subprocess.run(["make", "-s", "-C", str(HERE), str(HERE / "libhostemu_asan.so")], capture_output=True, text=True)
    asan = subprocess.run(["/usr/bin/g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if r.returncode != 0 or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("no sanitizer runtime for /usr/bin/g++")
    env = dict(os.environ, HOSTEMU_LIB=str(HERE / "libhostemu_asan.so"), LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0",
               UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([sys.executable, "-m", "pytest", str(pathlib.Path(__file__)), "-x", "-q", "-p", "no:cacheprovider"], capture_output=True, text=True,
                       timeout=1500, env=env)
    assert r.returncode == 0 and "Sanitizer" not in r.stderr and "runtime error" not in r.stderr, (r.stdout[-1500:], r.stderr[-3000:])
    assert " passed" in r.stdout and "1 skipped" in r.stdout  # the inner run skips this test
