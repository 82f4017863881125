"""The device-side sharded single-image path (BASELINE configs[3]) against the oracle.

On one GPU the ranks are virtual: several contexts (streams) of this process, control blocks shared by pointer.  With two
or more GPUs a second test runs one process per GPU with CUDA IPC mappings and leaves a log under profiles/."""
import importlib
import os
import pathlib
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def _frame(synth, w, h, first=0):
    return synth.iyuv_frames_numpy(w, h, 1, first)[0]


@pytest.mark.parametrize("w,h,world,q,full_frame", [(256, 112, 3, (50, 50, 50), False), (512, 256, 4, (90, 40, 75), True),
                                                    (64, 48, 4, (50, 50, 50), False), (1920, 1088, 2, (50, 50, 50), False),
                                                    (256, 160, 1, (10, 10, 10), True)])
def test_virtual_ranks_match_oracle(pkg, ora, synth, w, h, world, q, full_frame):
    torch = pytest.importorskip("torch")
    sharding = importlib.import_module("yuv-manipulations-2_b200.sharding")
    f = _frame(synth, w, h, 3)
    if w == 512:  # a noisy stripe: deferred blocks (more than 15 symbols) inside some bands
        rng = np.random.default_rng(3)
        f = f.copy()
        f[: w * h].reshape(h, w)[40:120, 100:300] = rng.integers(0, 256, (80, 200), dtype=np.uint8)
    want = ora.compress(f, w, h, q)
    ctxs = [pkg.Context(0) for _ in range(world)]
    groups = sharding.ShardGroup.local(ctxs, w, h)
    try:
        d_full = torch.from_numpy(f).cuda()
        bands = []
        for g in groups:
            y0, y1 = g.band
            b = torch.from_numpy(sharding.slice_iyuv(f, w, h, y0, y1)).cuda() if y1 > y0 else torch.zeros(8, dtype=torch.uint8, device="cuda")
            bands.append(b)
            if y1 > y0:
                # Allocate every context's workspace before the ranks have to meet on the device: growing a buffer frees the
                # old one, cudaFree waits for the device, and on ONE device that includes the other virtual ranks' waiting
                # kernels.  (One process per GPU has no such coupling.)
                bh = y1 - y0
                cap = pkg.capi.compress_bound(w, bh)
                t_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
                t_off = torch.zeros(2, dtype=torch.int64, device="cuda")
                t_back = torch.empty(w * bh * 3 // 2, dtype=torch.uint8, device="cuda")
                torch.cuda.synchronize()
                g.ctx.compress_batch_dev(b, w, bh, q, 1, t_out, cap, t_off)
                g.ctx.decompress_batch_dev(t_out, t_off, w, bh, q, 1, t_back)
                g.ctx.batch_status()
                # ... and the context's input buffer, where the decoding side keeps its copy of the band's payload (256 bytes per block)
                g.ctx.xrgb_to_iyuv(np.zeros(w * 2 * bh * 4, np.uint8), w, 2 * bh)
        torch.cuda.synchronize()
        for rep in range(3):  # epochs 1..3 on the same buffers
            for g, b in zip(groups, bands):
                g.compress(d_full if full_frame else b, q, full_frame)
            size = groups[0].result()
            for g in groups[1:]:
                g.ctx.batch_status()
            got = _device_bytes(torch, groups[0].root_out, size)
            assert size == want.size and np.array_equal(got, want), f"epoch {rep + 1}"
        # decode: every rank its band, gathered in the root's frame
        back = [torch.empty(max((g.band[1] - g.band[0]) * w * 3 // 2, 8), dtype=torch.uint8, device="cuda") for g in groups]
        torch.cuda.synchronize()
        for g, b in zip(groups, back):
            g.decompress(size, q, b)
        assert groups[0].result() == w * h * 3 // 2
        for g in groups[1:]:
            g.ctx.batch_status()
        dec = _device_bytes(torch, groups[0].root_iyuv, w * h * 3 // 2)
        assert np.array_equal(dec, ora.decompress(want, w, h, q))
        for g, b in zip(groups, back):
            y0, y1 = g.band
            if y1 > y0:
                assert np.array_equal(b[: (y1 - y0) * w * 3 // 2].cpu().numpy(), sharding.slice_iyuv(dec, w, h, y0, y1))
    finally:
        for g in groups:
            g.close()
        for c in ctxs:
            c.close()


def _device_bytes(torch, ptr, n):
    """n bytes of device memory at a raw address -> numpy (cudaMemcpy through the runtime torch loaded)."""
    import ctypes

    out = np.empty(n, np.uint8)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    torch.cuda.synchronize()
    assert rt.cudaMemcpy(out.ctypes.data, ctypes.c_void_p(ptr), n, 2) == 0
    return out


def test_missing_rank_times_out_instead_of_hanging(pkg, synth):
    """Only one of two ranks makes the call: it must come back with MYYUVB_ERR_SHARD_TIMEOUT after about 2 s."""
    pytest.importorskip("torch")
    import torch

    sharding = importlib.import_module("yuv-manipulations-2_b200.sharding")
    w, h = 128, 64
    ctxs = [pkg.Context(0), pkg.Context(0)]
    groups = sharding.ShardGroup.local(ctxs, w, h)
    try:
        f = _frame(synth, w, h)
        y0, y1 = groups[0].band
        b = torch.from_numpy(sharding.slice_iyuv(f, w, h, y0, y1)).cuda()
        torch.cuda.synchronize()
        groups[0].compress(b, (50, 50, 50))
        with pytest.raises(pkg.MyyuvError, match="did not arrive"):
            groups[0].result()
    finally:
        for g in groups:
            g.close()
        for c in ctxs:
            c.close()


WORKER = r'''
import importlib, json, os, sys, time
root = sys.argv[1]
sys.path.insert(0, root)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
pkg = importlib.import_module("yuv-manipulations-2_b200")
synth = importlib.import_module("yuv-manipulations-2_b200.sharding")
sharding = synth
gen = importlib.import_module("yuv-manipulations-2_b200.synth")
import oracle
w, h, q = 7680, 4320, (50, 50, 50)
f = gen.iyuv_frames_numpy(w, h, 1, 1)[0]
ctx = pkg.Context(rank)
g = sharding.ShardGroup.distributed(ctx, dist, w, h)
y0, y1 = g.band
band = torch.from_numpy(sharding.slice_iyuv(f, w, h, y0, y1)).cuda()
back = torch.empty_like(band)
torch.cuda.synchronize(); dist.barrier()
g.compress(band, q)
size = g.result() if rank == 0 else (ctx.batch_status() or 0)
sz = torch.tensor([size], dtype=torch.int64, device="cuda"); dist.broadcast(sz, 0); size = int(sz.item())
g.decompress(size, q, back)
if rank == 0:
    assert g.result() == w * h * 3 // 2
else:
    ctx.batch_status()
dist.barrier()
if rank == 0:
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    got = np.empty(size, np.uint8); rt.cudaMemcpy(got.ctypes.data, ctypes.c_void_p(g.root_out), size, 2)
    dec = np.empty(w * h * 3 // 2, np.uint8); rt.cudaMemcpy(dec.ctypes.data, ctypes.c_void_p(g.root_iyuv), dec.size, 2)
    ora = oracle.Oracle()
    want = ora.compress(f, w, h, q)
    ok = size == want.size and np.array_equal(got, want)
    ok_dec = np.array_equal(dec, ora.decompress(want, w, h, q))
    print(json.dumps({"world": world, "w": w, "h": h, "payload_bytes": size, "payload_equals_oracle": bool(ok), "decoded_equals_oracle": bool(ok_dec)}), flush=True)
    assert ok and ok_dec
dist.barrier()
g.close(); ctx.close()
dist.destroy_process_group()
'''


def test_sharded_8k_image_over_all_gpus(tmp_path):
    """One process per GPU, CUDA IPC mappings, bands stored into rank 0's buffer over NVLink: a 7680x4320 frame must come out
    byte-identical to the oracle's single-image payload and decode to the oracle's image.  Leaves a log under profiles/."""
    torch = pytest.importorskip("torch")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script), str(ROOT)], env=env, capture_output=True, text=True, timeout=900)
    log = ROOT / "profiles" / f"r02_shard8k_test_{n}gpu.log"
    try:
        log.write_text(r.stdout[-4000:] + "\n--- stderr tail ---\n" + r.stderr[-3000:])
    except OSError:
        pass
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"payload_equals_oracle": true' in r.stdout and '"decoded_equals_oracle": true' in r.stdout
