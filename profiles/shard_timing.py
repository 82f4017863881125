#!/usr/bin/env python3
"""Host and device cost of one sharded call (run under torchrun or alone): per-call host time of the C-ABI call and the
device time between events around one call, for compress and decompress.  Diagnostic behind profiles/r02_notes.md."""
import importlib, json, os, pathlib, sys, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pkg = importlib.import_module("yuv-manipulations-2_b200")
synth = importlib.import_module("yuv-manipulations-2_b200.synth")
sharding = importlib.import_module("yuv-manipulations-2_b200.sharding")
w, h, q = 7680, 4320, (50, 50, 50)
f = synth.iyuv_frames_numpy(w, h, 1, 1)[0]
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = pkg.Context(rank, stream.cuda_stream)
g = sharding.ShardGroup.distributed(ctx, dist, w, h) if world > 1 else sharding.ShardGroup.local([ctx], w, h)[0]
y0, y1 = g.band
band = torch.from_numpy(sharding.slice_iyuv(f, w, h, y0, y1)).to(dev); back = torch.empty_like(band)
torch.cuda.synchronize()
def fin(): return g.result() if rank == 0 else (ctx.batch_status() or 0)
def sync_all():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
for _ in range(3): g.compress(band, q)
size = fin()
st = torch.tensor([size], dtype=torch.int64, device=dev)
if world > 1: dist.broadcast(st, 0)
size = int(st.item())
for _ in range(3): g.decompress(size, q, back)
fin()
out = {}
for name, fn in (("compress", lambda: g.compress(band, q)), ("decompress", lambda: g.decompress(size, q, back))):
    host, devt = [], []
    for _ in range(10):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); e1.record(stream)
        fin(); devt.append(e0.elapsed_time(e1) * 1e3); host.append((t1 - t0) * 1e6)
    out[name] = {"host_us_per_call": round(float(np.median(host)), 1), "device_us_one_call": round(float(np.median(devt)), 1), "library_events_us": round(ctx.last_kernel_ms() * 1e3, 1)}
print(json.dumps({"rank": rank, "world": world, **out}), flush=True)
sync_all(); g.close()
if world > 1: dist.destroy_process_group()
