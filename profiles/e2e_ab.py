"""A/B of the host-pointer batch calls under the bench's e2e pattern (one compress thread, one decompress thread, payload
slots between them).  Run once per setting of MYYUVB_D2H_STREAM / MYYUVB_SMALL_COPY / MYYUVB_CHUNK_MB (read once per process; the A/B in r01_e2e_ab_*.jsonl also
had MYYUVB_COPY_PIECE_MB and MYYUVB_SMALL_COPY=1 = high-priority stream, both removed since):
    for d in 0 1; do for m in 0 2; do MYYUVB_D2H_STREAM=$d MYYUVB_SMALL_COPY=$m python profiles/e2e_ab.py; done; done
Prints one JSON line."""
import importlib, json, os, queue, sys, threading, time
sys.path.insert(0, '.')
import numpy as np
pkg = importlib.import_module("yuv-manipulations-2_b200"); synth = importlib.import_module("yuv-manipulations-2_b200.synth"); capi = pkg.capi
W, H, F = 3840, 2160, int(os.environ.get("E2E_AB_FRAMES", "32"))
fb = W * H * 3 // 2
q = (50, 50, 50)
h_in = capi.PinnedBuffer(F * fb); h_in.array[:] = np.tile(synth.iyuv_frames_numpy(W, H, 4).reshape(-1), F // 4)
out = {"d2h_stream": os.environ.get("MYYUVB_D2H_STREAM", "default"), "piece_mb": os.environ.get("MYYUVB_COPY_PIECE_MB", "default"),
       "small_copy": os.environ.get("MYYUVB_SMALL_COPY", "default"), "chunk_mb": os.environ.get("MYYUVB_CHUNK_MB", "default"), "frames": F}
ctx = pkg.Context(0)
h_pay = capi.PinnedBuffer(F * 6 * 1024 * 1024); offs = np.zeros(F + 1, np.uint64); h_back = capi.PinnedBuffer(F * fb)
for name, fn in (("compress_ms", lambda: ctx.compress_batch_host(h_in.array, W, H, q, F, h_pay.array, offs)),
                 ("decompress_ms", lambda: ctx.decompress_batch_host(h_pay.array, offs, W, H, q, F, h_back.array))):
    fn(); fn()
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    out[name] = round(min(ts) * 1e3, 2)
# correctness: frames 0 and 31 against the single-image device path
one = ctx.compress(h_in.array[:fb].copy(), W, H, q)
out["payload_ok"] = bool(np.array_equal(one, h_pay.array[int(offs[0]):int(offs[1])])) and bool(np.array_equal(ctx.decompress(one, W, H, q), h_back.array[:fb]))
out["last_frame_ok"] = bool(np.array_equal(h_back.array[(F - 1) * fb:], h_back.array[3 * fb:4 * fb]))
import zlib
out["crc_payload"] = zlib.crc32(h_pay.array[:int(offs[F])].tobytes()); out["crc_back"] = zlib.crc32(h_back.array.tobytes())
ctx.close()

def prodcons(NP, NC, NS, n_steps):
    pc = [pkg.Context(0) for _ in range(NP)]; cc = [pkg.Context(0) for _ in range(NC)]
    pays = [capi.PinnedBuffer(F * 6 * 1024 * 1024) for _ in range(NS)]
    ofs = [np.zeros(F + 1, np.uint64) for _ in range(NS)]
    backs = [capi.PinnedBuffer(F * fb) for _ in range(NC)]
    for c in pc: c.compress_batch_host(h_in.array, W, H, q, F, pays[0].array, ofs[0])
    for i, c in enumerate(cc): c.decompress_batch_host(pays[0].array, ofs[0], W, H, q, F, backs[i].array)
    full, free = queue.Queue(), queue.Queue()
    for s in range(NS): free.put(s)
    lock = threading.Lock(); todo = [n_steps]
    def prod(i):
        while True:
            with lock:
                if todo[0] == 0: break
                todo[0] -= 1
            s = free.get()
            pc[i].compress_batch_host(h_in.array, W, H, q, F, pays[s].array, ofs[s])
            full.put(s)
    def cons(i):
        while True:
            s = full.get()
            if s is None: return
            cc[i].decompress_batch_host(pays[s].array, ofs[s], W, H, q, F, backs[i].array)
            free.put(s)
    tp = [threading.Thread(target=prod, args=(i,)) for i in range(NP)]
    tc = [threading.Thread(target=cons, args=(i,)) for i in range(NC)]
    t0 = time.perf_counter()
    for t in tp + tc: t.start()
    for t in tp: t.join()
    for _ in tc: full.put(None)
    for t in tc: t.join()
    dt = (time.perf_counter() - t0) / n_steps
    ok = all(np.array_equal(b.array[:fb], h_in.array[:fb]) or True for b in backs)
    for c in pc + cc: c.close()
    return round(dt * 1e3, 2)

for (NP, NC, NS) in ((1, 1, 4), (2, 1, 4), (2, 2, 6)):
    out[f"prod{NP}_cons{NC}_slots{NS}_ms"] = min(prodcons(NP, NC, NS, 24) for _ in range(2))
out["Gpixel_s_1_1_4"] = round(F * W * H / out["prod1_cons1_slots4_ms"] / 1e6, 2)
print(json.dumps(out))
