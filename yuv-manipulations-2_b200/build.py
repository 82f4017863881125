"""Build the native parts of the package in-tree (so the .so files travel with the repo snapshot).

  lib/libmyyuvb200.so   CUDA kernels + C ABI (include/myyuvb200.h), sm_100a only
  lib/libmyyuv_lib.so   drop-in C++ class library (myyuv::BMP / myyuv::YUV, include/myyuv*.hpp) on top of the C ABI
  lib/myyuv_cli         the reference's UNMODIFIED CLI object (oracle/_ref/myyuv_cli_main.o) linked against it, when present

Flags that matter for bit-exactness: -fmad=false for device code (no FMUL+FADD -> FFMA contraction) and
-ffp-contract=off for host code (quantisation tables use the reference's float expression).
"""
from __future__ import annotations

import os
import pathlib
import shutil
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("MYYUVB_CXX", "/usr/bin/g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: pathlib.Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(pathlib.Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True)


def build_cuda(force: bool = False, verbose: bool = False, variant: str = "") -> pathlib.Path:
    """variant "clk": the same sources with -DMYYUVB_PHASE_CLOCKS (per-phase clock counters in the codec kernels), a
    profiling build next to the product library, loaded only when MYYUVB_LIB_VARIANT=clk (profiles/phase_clocks.py).
    variant "notma": -DMYYUVB_NO_TMA_STAGE, the decoder's chunk staging as a loop of 128-bit loads instead of the 1-D bulk
    copy (the A/B in profiles/r02_notes.md)."""
    LIB.mkdir(exist_ok=True)
    out = LIB / ("libmyyuvb200.so" if not variant else f"libmyyuvb200_{variant}.so")
    srcs = [CSRC / "kernels.cu", CSRC / "capi.cu"]
    deps = srcs + [CSRC / "kernels.h", CSRC / "block_codec.cuh", CSRC / "dct_matrix.inc", ROOT / "include/myyuvb200.h"]
    if force or _newer(out, deps):
        extra = {"clk": ["-DMYYUVB_PHASE_CLOCKS"], "notma": ["-DMYYUVB_NO_TMA_STAGE"]}.get(variant, [])
        extra += os.environ.get("MYYUVB_VARIANT_FLAGS", "").split() if variant else []  # experiments: profiles/ab.py
        _run([NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-fmad=false", *extra, "-ccbin", CXX,
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "--shared", "-o", out, *srcs], verbose)
    return out


def device_code_sha256(lib: pathlib.Path | None = None) -> str | None:
    """sha256 of the library's SASS (cuobjdump -sass, comment and blank lines dropped): the identity of the machine code the
    GPU runs.  profiles/launch_list.py records it with an ncu capture and bench.py reports the capture's DRAM traffic only when
    the library it has loaded hashes the same -- an edit of a comment keeps a capture valid, another compiler flag does not.
    None when cuobjdump is not available."""
    import hashlib

    lib = lib or (LIB / "libmyyuvb200.so")
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        r = subprocess.run([tool, "-sass", str(lib)], capture_output=True, text=True, timeout=120)
    except (OSError, subprocess.SubprocessError):
        return None
    if r.returncode != 0 or "Function :" not in r.stdout:
        return None
    h = hashlib.sha256()
    for line in r.stdout.splitlines():
        t = line.strip()
        if t and not t.startswith("//"):
            h.update(t.encode() + b"\n")
    return h.hexdigest()


def build_cxx(force: bool = False, verbose: bool = False):
    """The drop-in class library and (if the reference CLI object is available) the unmodified CLI on top of it."""
    LIB.mkdir(exist_ok=True)
    out = LIB / "libmyyuv_lib.so"
    srcs = [CSRC / "bmp_host.cpp", CSRC / "yuv_host.cpp"]
    if not all(s.exists() for s in srcs):
        return None
    deps = srcs + list((ROOT / "include").glob("*.h*"))
    if force or _newer(out, deps):
        _run([CXX, "-std=gnu++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", f"-I{ROOT / 'include'}", *srcs,
              f"-L{LIB}", "-lmyyuvb200", "-Wl,-rpath,$ORIGIN", "-o", out], verbose)
    cli_obj = ROOT / "oracle/_ref/myyuv_cli_main.o"
    cli = LIB / "myyuv_cli"
    if cli_obj.exists() and (force or _newer(cli, [cli_obj, out])):
        _run([CXX, cli_obj, f"-L{LIB}", "-lmyyuv_lib", "-lmyyuvb200", "-Wl,-rpath,$ORIGIN", "-o", cli], verbose)
    return out


def build(force: bool = False, verbose: bool = False):
    build_cuda(force, verbose)
    build_cxx(force, verbose)


if __name__ == "__main__":
    if "--variant" in sys.argv:  # build.py --variant NAME  with MYYUVB_VARIANT_FLAGS="-D..." in the environment
        build_cuda(force=True, verbose=True, variant=sys.argv[sys.argv.index("--variant") + 1])
    elif "--clk" in sys.argv or "--notma" in sys.argv:
        build_cuda(force="--force" in sys.argv, verbose=True, variant="clk" if "--clk" in sys.argv else "notma")
    else:
        build(force="--force" in sys.argv, verbose=True)
