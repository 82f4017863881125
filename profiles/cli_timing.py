#!/usr/bin/env python3
"""Wall clock of the reference's UNMODIFIED CLI, one process per operation (what a CLI user waits for), three ways:
the reference library (serial and OpenMP builds) and the drop-in library (lib/myyuv_cli = the same main.o linked against
lib/libmyyuv_lib.so -> sm_100a kernels).  Sizes: the sample image, a 4K and an 8K synthetic frame.
   python profiles/cli_timing.py > profiles/r02_cli_timing.json"""
import importlib, json, os, pathlib, re, subprocess, sys, tempfile, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("yuv-manipulations-2_b200")
synth = importlib.import_module("yuv-manipulations-2_b200.synth")
CLIS = {"reference serial": ROOT / "oracle/_ref/serial/myyuv_cli", "reference OpenMP": ROOT / "oracle/_ref/omp/myyuv_cli",
        "drop-in (B200)": ROOT / "yuv-manipulations-2_b200/lib/myyuv_cli"}
tmp = pathlib.Path(tempfile.mkdtemp())
out = []
for w, h in ((992, 736), (3840, 2160), (7680, 4320)):
    y = pkg.YUV()
    y.header.fourcc_format = pkg.YUV.FourccFormats.IYUV
    y.header.width, y.header.height, y.header.data_size, y.header.data_pos = w, h, w * h * 3 // 2, 64
    y.data = synth.iyuv_frames_numpy(w, h, 1, 0)[0]
    src = tmp / f"src_{w}.myyuv"
    y.dump(str(src))
    row = {"width": w, "height": h}
    files = {}
    for name, exe in CLIS.items():
        if not exe.exists():
            continue
        rec = {}
        for op, args, inp in (("compress", ["-compress", "DCT", "50"], src), ("decompress", ["-decompress"], None)):
            inp = inp or files[name]
            o = tmp / f"{op}_{w}_{name.split()[0]}_{name.split()[-1]}.myyuv"
            best, own = None, None
            for _ in range(3):
                t0 = time.perf_counter()
                r = subprocess.run([str(exe), str(inp), *args, "-o", str(o)], capture_output=True, text=True)
                dt = 1e3 * (time.perf_counter() - t0)
                assert r.returncode == 0 and "Success!" in r.stdout, r.stdout + r.stderr
                m = re.search(r":\s*(\d+)\s*ms", r.stdout)
                if best is None or dt < best:
                    best, own = dt, int(m.group(1)) if m else None
            rec[op] = {"process_wall_ms": round(best, 1), "cli_timer_ms": own}
            if op == "compress":
                files[name] = o
        row[name] = rec
    a = (tmp / f"compress_{w}_reference_serial.myyuv")
    b = (tmp / f"compress_{w}_drop-in_(B200).myyuv")
    if a.exists() and b.exists():
        row["identical_files"] = a.read_bytes() == b.read_bytes()
    out.append(row)
print(json.dumps({"threads": os.cpu_count(), "note": "best of 3 process launches; cli_timer_ms = the CLI's own MyTimer print (integer ms around the operation only)", "rows": out}, indent=1))
