#!/usr/bin/env python3
"""Regenerate tests/golden/golden.json and the small binary fixtures FROM THE REFERENCE ITSELF.

Run in the build container (needs oracle/_ref, i.e. /root/reference compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
Inputs are the deterministic synthetic frames of yuv-manipulations-2_b200/synth.py (regenerated from the seed
wherever the tests run) plus the reference's own sample images; outputs are produced by the UNMODIFIED
reference library (serial build) through oracle/ref_shim.cpp.  Only hashes and two tiny fixtures are stored.
"""
import hashlib
import importlib
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402

synth = importlib.import_module("yuv-manipulations-2_b200.synth")
HERE = pathlib.Path(__file__).resolve().parent


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = O.Reference("serial")
    out = {"generator": "tests/golden/make_golden.py", "reference": "mahbhlddnhakkh/yuv-manipulations-2 (oracle/_ref/serial)",
           "synthetic": [], "colour": [], "edge": [], "chef": {}}
    cases = [(16, 16, (50, 50, 50), 0), (48, 80, (50, 50, 50), 1), (256, 128, (10, 50, 90), 2), (256, 128, (100, 100, 100), 3),
             (992, 736, (50, 50, 50), 4), (992, 736, (90, 90, 90), 5), (1920, 1088, (50, 50, 50), 6), (1920, 1088, (1, 25, 75), 7),
             (3840, 2160, (50, 50, 50), 0)]
    for w, h, q, first in cases:
        f = synth.iyuv_frames_numpy(w, h, 1, first)[0]
        c = ref.compress(f, w, h, q)
        d = ref.decompress(c, w, h, q)
        out["synthetic"].append({"w": w, "h": h, "q": list(q), "first": first, "input_sha256": sha(f), "payload_size": int(c.size),
                                 "payload_sha256": sha(c), "decoded_sha256": sha(d)})
    for w, h, first in [(16, 16, 0), (64, 48, 1), (992, 736, 2), (1920, 1080, 3)]:
        b = synth.bgrx_frames_numpy(w, h, 1, first)[0]
        for bottom_up in (True, False):
            y = ref.bgrx_to_iyuv(b, w, h, bottom_up)
            out["colour"].append({"w": w, "h": h, "first": first, "bottom_up": bottom_up, "input_sha256": sha(b), "iyuv_sha256": sha(y)})
    # 24-bit BMP rows (the reference's Release build converts them as B,G,R triplets): the same frames without the X byte
    out["colour24"] = []
    for w, h, first in [(16, 16, 0), (64, 48, 1), (36, 20, 5), (992, 736, 2), (1920, 1080, 3)]:
        b = np.ascontiguousarray(synth.bgrx_frames_numpy(w, h, 1, first)[0].reshape(-1, 4)[:, :3]).reshape(-1)
        for bottom_up in (True, False):
            y = ref.bgr24_to_iyuv(b, w, h, bottom_up)
            out["colour24"].append({"w": w, "h": h, "first": first, "bottom_up": bottom_up, "input_sha256": sha(b), "iyuv_sha256": sha(y)})
    for q in [(50, 50, 50), (1, 1, 1), (100, 100, 100), (97, 3, 64)]:
        f = synth.edge_case_iyuv(128, 128)
        c = ref.compress(f, 128, 128, q)
        out["edge"].append({"w": 128, "h": 128, "q": list(q), "input_sha256": sha(f), "payload_size": int(c.size), "payload_sha256": sha(c),
                            "decoded_sha256": sha(ref.decompress(c, 128, 128, q))})
    # two tiny fixtures kept as bytes so a failure can be inspected without the reference
    f = synth.iyuv_frames_numpy(32, 32, 1, 9)[0]
    c = ref.compress(f, 32, 32, (50, 50, 50))
    np.save(HERE / "tiny_32x32_q50_input.npy", f)
    np.save(HERE / "tiny_32x32_q50_payload.npy", c)
    np.save(HERE / "tiny_32x32_q50_decoded.npy", ref.decompress(c, 32, 32, (50, 50, 50)))
    # random blocks through Huffman::fromData/dump (all rehash steps of the tie-break emulation)
    rng = np.random.default_rng(20261018)
    blocks = np.zeros((4096, 64), np.int16)
    for i in range(4096):
        m = int(rng.integers(1, 65))
        vals = rng.choice(np.arange(-1024, 1024), m, replace=False)
        blocks[i] = vals[rng.integers(0, m, 64)]
        if i % 3 == 0:
            blocks[i, rng.integers(0, 64, 40)] = 0
    chunks, sizes = ref.huff_encode_blocks(blocks)
    out["huffman_blocks"] = {"seed": 20261018, "n": 4096, "sizes_sha256": sha(sizes), "chunks_sha256": sha(chunks)}
    # natural content at the benchmark's frame size (SURVEY 8(d) config 3(i) "tiled-real"): the reference's sample image
    # tiled to 3840x2160, origin shifted per frame
    out["tiled_real"] = []
    if (O.GOLDEN_DIR / "chef-with-trumpet.myyuv").exists():
        g = O.read_myyuv(O.GOLDEN_DIR / "chef-with-trumpet.myyuv")
        for q, first in [((50, 50, 50), 0), ((90, 90, 90), 3), ((10, 10, 10), 5)]:
            f = synth.tiled_real_iyuv(g["data"], g["w"], g["h"], 3840, 2160, 1, first)[0]
            c = ref.compress(f, 3840, 2160, q)
            out["tiled_real"].append({"w": 3840, "h": 2160, "q": list(q), "first": first, "input_sha256": sha(f), "payload_size": int(c.size),
                                      "payload_sha256": sha(c), "decoded_sha256": sha(ref.decompress(c, 3840, 2160, q))})
        # BASELINE configs[3]: one 7680x4320 frame (the image the sharded path codes); also its synthetic twin
        out["shard8k"] = []
        for content, first in (("tiled-real", 0), ("noise-grad", 1)):
            f = (synth.tiled_real_iyuv(g["data"], g["w"], g["h"], 7680, 4320, 1, first)[0] if content == "tiled-real"
                 else synth.iyuv_frames_numpy(7680, 4320, 1, first)[0])
            c = ref.compress(f, 7680, 4320, (50, 50, 50))
            out["shard8k"].append({"w": 7680, "h": 4320, "q": [50, 50, 50], "content": content, "first": first, "input_sha256": sha(f),
                                   "payload_size": int(c.size), "payload_sha256": sha(c),
                                   "decoded_sha256": sha(ref.decompress(c, 7680, 4320, (50, 50, 50)))})
    # the reference's sample images
    if (O.GOLDEN_DIR / "chef-with-trumpet.bmp").exists():
        for name in ["chef-with-trumpet.bmp", "chef-with-trumpet.myyuv", "chef-with-trumpet-DCT-50.myyuv", "chef-with-trumpet-DCT-90.myyuv",
                     "chef-with-trumpet-big-DCT-50.myyuv"]:
            out["chef"][name] = hashlib.sha256((O.GOLDEN_DIR / name).read_bytes()).hexdigest()
        for name in ["chef-with-trumpet-DCT-50.myyuv", "chef-with-trumpet-DCT-90.myyuv", "chef-with-trumpet-big-DCT-50.myyuv"]:
            g = O.read_myyuv(O.GOLDEN_DIR / name)
            out["chef"]["decoded:" + name] = sha(ref.decompress(g["data"], g["w"], g["h"], g["params"]))
    (HERE / "golden.json").write_text(json.dumps(out, indent=1))
    print("wrote", HERE / "golden.json")


if __name__ == "__main__":
    main()
