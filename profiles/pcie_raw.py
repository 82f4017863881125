"""Raw pinned-memory copy rates of the box: each direction alone, both at once (one thread, two streams), and both at once
in 62 MB pieces from two host threads (the shape of the batch_host calls).  One JSON line."""
import json, threading, time
import torch
N = 32 * 3840 * 2160 * 3 // 2
d1 = torch.empty(N, dtype=torch.uint8, device="cuda"); d2 = torch.empty(N, dtype=torch.uint8, device="cuda")
h1 = torch.empty(N, dtype=torch.uint8).pin_memory(); h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
out = {}
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
out["h2d_GBps"] = round(N / timed(h2d) / 1e9, 1)
out["d2h_GBps"] = round(N / timed(d2h) / 1e9, 1)
out["both_total_GBps"] = round(2 * N / timed(lambda: (h2d(), d2h())) / 1e9, 1)
P = 5 * 3840 * 2160 * 3 // 2
def pieces(dst, src, stream, sync_each):
    with torch.cuda.stream(stream):
        for o in range(0, N, P):
            dst[o:o + P].copy_(src[o:o + P], non_blocking=True)
            if sync_each: stream.synchronize()
        stream.synchronize()
for sync_each in (False, True):
    def both():
        ta = threading.Thread(target=pieces, args=(d1, h1, s1, sync_each)); tb = threading.Thread(target=pieces, args=(h2, d2, s2, sync_each))
        ta.start(); tb.start(); ta.join(); tb.join()
    both()
    t0 = time.perf_counter()
    for _ in range(5): both()
    dt = (time.perf_counter() - t0) / 5
    out[f"two_threads_pieces_sync{int(sync_each)}_total_GBps"] = round(2 * N / dt / 1e9, 1)
print(json.dumps(out))
