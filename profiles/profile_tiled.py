"""ncu target: one compress + decompress of 8 natural-content 4K frames (SURVEY 8(d)(i) tiled-real) at the quality in argv[1]."""
import importlib, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
pkg = importlib.import_module("yuv-manipulations-2_b200"); synth = importlib.import_module("yuv-manipulations-2_b200.synth")
q = int(sys.argv[1]) if len(sys.argv) > 1 else 50
W, H, FR = 3840, 2160, 8
raw = (ROOT / "oracle" / "_ref" / "golden" / "chef-with-trumpet.myyuv").read_bytes()
base = np.frombuffer(raw, np.uint8, 992 * 736 * 3 // 2, 64).copy()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = pkg.Context(0, stream.cuda_stream)
r_in = torch.from_numpy(synth.tiled_real_iyuv(base, 992, 736, W, H, FR)).to(dev)
r_out = torch.empty(FR * pkg.capi.compress_bound(W, H), dtype=torch.uint8, device=dev)
r_off = torch.zeros(FR + 1, dtype=torch.int64, device=dev)
r_back = torch.empty_like(r_in)
for _ in range(2):
    ctx.compress_batch_dev(r_in, W, H, (q, q, q), FR, r_out, r_out.numel(), r_off)
    ctx.decompress_batch_dev(r_out, r_off, W, H, (q, q, q), FR, r_back)
ctx.batch_status()
print("ok", int(r_off[FR].item()))
