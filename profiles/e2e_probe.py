import importlib, sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
pkg = importlib.import_module("yuv-manipulations-2_b200"); synth = importlib.import_module("yuv-manipulations-2_b200.synth"); capi = pkg.capi
W,H,F=3840,2160,32
fb=W*H*3//2
q=(50,50,50)
ctx=pkg.Context(0)
h_in=capi.PinnedBuffer(F*fb); h_in.array[:]=synth.iyuv_frames_numpy(W,H,4).reshape(-1).repeat(1)[:4*fb].tolist() if False else np.tile(synth.iyuv_frames_numpy(W,H,4).reshape(-1), F//4)
h_pay=capi.PinnedBuffer(F*6*1024*1024); offs=np.zeros(F+1,np.uint64); h_back=capi.PinnedBuffer(F*fb)
out={}
for name,fn in (("compress",lambda: ctx.compress_batch_host(h_in.array,W,H,q,F,h_pay.array,offs)),("decompress",lambda: ctx.decompress_batch_host(h_pay.array,offs,W,H,q,F,h_back.array))):
    fn(); fn()
    t0=time.perf_counter()
    for _ in range(5): fn()
    dt=(time.perf_counter()-t0)/5
    out[name]={"ms":round(dt*1e3,2)}
pay=int(offs[F]); out["payload"]=pay
out["compress"]["GBps_h2d"]=round(F*fb/out["compress"]["ms"]/1e6,1)
out["decompress"]["GBps_d2h"]=round(F*fb/out["decompress"]["ms"]/1e6,1)
# raw pinned copies
d=torch.empty(F*fb,dtype=torch.uint8,device="cuda")
hin=torch.from_numpy(h_in.array); hb=torch.from_numpy(h_back.array)
for nm,(a,b) in (("h2d",(d,hin)),("d2h",(hb,d))):
    a.copy_(b,non_blocking=True); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(5): a.copy_(b,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5
    out["raw_"+nm+"_GBps"]=round(F*fb/dt/1e9,1)
print(json.dumps(out))
# raw concurrent H2D + D2H on two streams (what the platform allows when both directions are busy)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(F * fb, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2):
        hb.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(json.dumps({"raw_bidirectional_total_GBps": round(2 * F * fb / dt / 1e9, 1)}))
# the two batch_host calls from two host threads (two contexts), as bench.py's e2e leg runs them
import threading
c2 = pkg.Context(0)
offs2 = offs.copy()
h_pay2 = capi.PinnedBuffer(F * 6 * 1024 * 1024)
c2.compress_batch_host(h_in.array, W, H, q, F, h_pay2.array, offs2)
def run_c(n):
    for _ in range(n): ctx.compress_batch_host(h_in.array, W, H, q, F, h_pay.array, offs)
def run_d(n):
    for _ in range(n): c2.decompress_batch_host(h_pay2.array, offs2, W, H, q, F, h_back.array)
for rep in range(2):
    t0 = time.perf_counter()
    ta = threading.Thread(target=run_c, args=(5,)); tb = threading.Thread(target=run_d, args=(5,))
    ta.start(); tb.start(); ta.join(); tb.join()
    dt = (time.perf_counter() - t0) / 5
print(json.dumps({"concurrent_compress_and_decompress_ms_per_pair": round(dt * 1e3, 2)}))
# W round-trip workers (compress then decompress, own context and buffers each)
import queue
def roundtrip_workers(NW, n_steps, stagger):
    cs = [pkg.Context(0) for _ in range(NW)]
    pays = [capi.PinnedBuffer(F * 6 * 1024 * 1024) for _ in range(NW)]
    ofs = [np.zeros(F + 1, np.uint64) for _ in range(NW)]
    backs = [capi.PinnedBuffer(F * fb) for _ in range(NW)]
    for i in range(NW):
        cs[i].compress_batch_host(h_in.array, W, H, q, F, pays[i].array, ofs[i])
        cs[i].decompress_batch_host(pays[i].array, ofs[i], W, H, q, F, backs[i].array)
    lock = threading.Lock(); todo = [n_steps]
    def worker(i):
        while True:
            with lock:
                if todo[0] == 0: return
                todo[0] -= 1
            if stagger and (i & 1):
                cs[i].decompress_batch_host(pays[i].array, ofs[i], W, H, q, F, backs[i].array)
                cs[i].compress_batch_host(h_in.array, W, H, q, F, pays[i].array, ofs[i])
            else:
                cs[i].compress_batch_host(h_in.array, W, H, q, F, pays[i].array, ofs[i])
                cs[i].decompress_batch_host(pays[i].array, ofs[i], W, H, q, F, backs[i].array)
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(NW)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    dt = (time.perf_counter() - t0) / n_steps
    for c in cs: c.close()
    return round(dt * 1e3, 2)
res = {}
for NW in (1, 2, 3, 4, 6):
    for st in (0, 1):
        res[f"workers{NW}_stagger{st}"] = roundtrip_workers(NW, 12, st)
print(json.dumps(res))
# dedicated compress threads -> queue of payload slots -> dedicated decompress threads
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
    for c in pc + cc: c.close()
    return round(dt * 1e3, 2)
res = {}
for (NP, NC, NS) in ((1, 1, 2), (1, 1, 4), (2, 2, 4), (2, 2, 6), (1, 2, 4), (2, 3, 6)):
    res[f"prod{NP}_cons{NC}_slots{NS}"] = prodcons(NP, NC, NS, 16)
print(json.dumps(res))
