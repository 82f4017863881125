"""Host logic of the multi-GPU path (yuv-manipulations-2_b200/sharding.py) on CPU: byte-level assembly/splitting of
band payloads, and the world_size-2 drivers over torch.distributed with the gloo backend.  The coder plugged in
here is the CPU oracle (the module only moves bytes); on GPUs the same drivers get the CUDA context's methods."""
import importlib
import os
import pathlib
import subprocess
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sh = importlib.import_module("yuv-manipulations-2_b200.sharding")


def test_frame_ranges_cover_batch():
    for n in (1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            r = [sh.frame_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_macroblock_bands_8k():
    bands = sh.macroblock_row_bands(4320, 8)
    assert [(b - a) // 16 for a, b in bands] == [34, 34, 34, 34, 34, 34, 33, 33]
    assert bands[0][0] == 0 and bands[-1][1] == 4320
    assert sh.macroblock_row_bands(32, 4) == [(0, 16), (16, 32), (32, 32), (32, 32)]
    with pytest.raises(ValueError):
        sh.macroblock_row_bands(40, 2)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_band_payloads_assemble_to_whole_image_payload(ora, synth, world):
    w, h, q = 128, 96 + 16 * world, (50, 70, 30)
    img = synth.iyuv_frames_numpy(w, h, 1, 4)[0]
    whole = ora.compress(img, w, h, q)
    ranges = sh.macroblock_row_bands(h, world)
    bands = [ora.compress(sh.slice_iyuv(img, w, h, a, b), w, b - a, q) if b > a else np.zeros(0, np.uint8) for a, b in ranges]
    assert np.array_equal(sh.assemble_payload(bands), whole)
    # and back: splitting the whole payload gives exactly the band payloads
    for got, want in zip(sh.split_payload(whole, w, ranges), bands):
        assert np.array_equal(got, want)
    dec = [ora.decompress(b, w, y1 - y0, q) if y1 > y0 else np.zeros(0, np.uint8) for b, (y0, y1) in zip(bands, ranges)]
    assert np.array_equal(sh.unslice_iyuv(dec, w, h, ranges), ora.decompress(whole, w, h, q))


def test_parse_build_round_trip(ora, synth):
    img = synth.iyuv_frames_numpy(64, 64, 1)[0]
    p = ora.compress(img, 64, 64, (90, 90, 90))
    planes = sh.parse_payload(p)
    assert [n for n, _, _ in planes] == [64, 16, 16]
    assert np.array_equal(sh.build_payload([(s, c) for _, s, c in planes]), p)


WORKER = r'''
import importlib, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
import oracle
sh = importlib.import_module("yuv-manipulations-2_b200.sharding")
synth = importlib.import_module("yuv-manipulations-2_b200.synth")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ora = oracle.Oracle(threads=1)
w, h, q = 192, 160, (50, 50, 50)
img = synth.iyuv_frames_numpy(w, h, 1, 3)[0]           # every rank can regenerate the image; it only uses its band
y0, y1 = sh.macroblock_row_bands(h, world)[rank]
band = sh.slice_iyuv(img, w, h, y0, y1)
payload = sh.compress_image_sharded(band, w, y1 - y0, q, ora.compress, dist)
ok = True
if rank == 0:
    ok = np.array_equal(payload, ora.compress(img, w, h, q))
    assert ok, "assembled payload differs from the single-process payload"
dec = sh.decompress_image_sharded(payload if rank == 0 else None, w, h, q, ora.decompress, dist)
if rank == 0:
    assert np.array_equal(dec, ora.decompress(payload, w, h, q)), "sharded decode differs"
# batch sharding: each rank codes its own frame range, nothing is exchanged; sizes are summed only to check coverage
lo, hi = sh.frame_range(5, rank, world)
import torch
n = torch.tensor([hi - lo]); dist.all_reduce(n); assert int(n.item()) == 5
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank} ok")
'''


def test_sharded_image_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=str(ROOT), OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


GPU_WORKER = r'''
import importlib, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
import oracle
pkg = importlib.import_module("yuv-manipulations-2_b200")
sh = importlib.import_module("yuv-manipulations-2_b200.sharding")
synth = importlib.import_module("yuv-manipulations-2_b200.synth")
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = pkg.Context(local)
w, h, q = 7680, 4320, (50, 50, 50)                    # BASELINE configs[3]: one 8K image, macroblock rows sharded
y0, y1 = sh.macroblock_row_bands(h, world)[rank]
img = synth.iyuv_frames_numpy(w, h, 1, 1)[0] if rank == 0 else None
# every rank generates only its band (same integer formula, rows offset)
full = synth.iyuv_frames_numpy(w, h, 1, 1)[0] if img is None else img
band = sh.slice_iyuv(full, w, h, y0, y1)
payload = sh.compress_image_sharded(band, w, y1 - y0, q, ctx.compress, dist, device=f"cuda:{local}")
if rank == 0:
    ora = oracle.Oracle()
    want = ora.compress(full, w, h, q)
    assert np.array_equal(payload, want), "sharded 8K payload differs from the oracle's single-image payload"
dec = sh.decompress_image_sharded(payload if rank == 0 else None, w, h, q, ctx.decompress, dist, device=f"cuda:{local}")
if rank == 0:
    assert np.array_equal(dec, ora.decompress(want, w, h, q)), "sharded 8K decode differs"
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank} ok")
'''


@pytest.mark.gpu
def test_sharded_8k_image_nccl(tmp_path):
    torch = pytest.importorskip("torch")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    n = min(n, 8)
    script = tmp_path / "worker.py"
    script.write_text(GPU_WORKER)
    env = dict(os.environ, REPO_ROOT=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29612", str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert all(f"rank {k} ok" in r.stdout for k in range(n))
