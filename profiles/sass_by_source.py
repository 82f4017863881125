#!/usr/bin/env python3
"""Where a kernel's executed instructions come from, by source line and by region of the source.

`ncu --set full --import-source on` records "Instructions Executed" and stall samples per SASS instruction; the report's own
source page shows them per SASS line only.  This joins them with the line table of the library (nvdisasm --print-line-info on
the cubin inside the .so, same instruction order) and sums per (file, line) and per named region (line ranges below).

  python profiles/sass_by_source.py gpurun_out/r02f_ng.ncu-rep dct_compress_kernel  [top lines, default 30]
  python profiles/sass_by_source.py gpurun_out/r02f_ng.ncu-rep dct_decompress_kernel

The report and the library must be the same machine code (the tool checks that both list the same opcodes in the same order);
line numbers are those of the library's line table, i.e. of the sources it was built from.
"""
import collections, csv, io, pathlib, re, subprocess, sys, tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent
LIB = ROOT / "yuv-manipulations-2_b200" / "lib" / "libmyyuvb200.so"
CSRC = ROOT / "yuv-manipulations-2_b200" / "csrc"

# A region = (name, file, regex of its first line, regex of the first line behind it); both searched in the current sources, the
# second one behind the first.  The kernels' phases are delimited by their PH(n) markers (the phase-clock build's probes).
def _phases(kernel_re, names):
    out, prev = [], kernel_re
    for k, nm in enumerate(names):
        out.append((f"kernel phase {k}: {nm}", "kernels.cu", prev, rf"PH\({k}\);", kernel_re))
        prev = rf"PH\({k}\);"
    return out


ENC, DEC = r"^\s+dct_compress_kernel\(const __grid_constant__", r"^\s+dct_decompress_kernel\(const __grid_constant__"
REGIONS = {
    "dct_compress_kernel": [
        ("packed f32x2 helpers: the DCT's products and sums, exact division, rounding constants", "kernels.cu", r"^struct __align__\(8\) f2", r"^MYB_D constexpr float dct_c"),
        ("fdct_quant_block: byte->float, loops, F2I, zigzag stores, message length", "kernels.cu", r"^MYB_D int fdct_quant_block", r"^// Two builds of the kernel"),
        ("huff_hist_smem (PTX histogram loop)", "kernels.cu", r"^MYB_D int huff_hist_smem", r"^struct EncParams"),
        ("fast plan (huff_fast_plan_n): list order of the reference's map, leaves, merges, table size", "block_codec.cuh", r"^MYB_HD FastPlan huff_fast_plan_n", r"^MYB_HD FastPlan huff_fast_plan\("),
        ("fast plan: heap sift / pop", "block_codec.cuh", r"^MYB_HD void heap32_sift_up", r"^MYB_HD void lst_place"),
        ("fast plan: nibble list / bucket helpers", "block_codec.cuh", r"^MYB_HD uint32_t spread8", r"^MYB_HD void heap_push_static"),
        ("FastScratch accessors (shared-memory addresses of slots, heap, code words)", "block_codec.cuh", r"^struct FastScratch", r"^MYB_HD int huff_hist"),
        ("fast emit (huff_fast_emit_n): sort, canonical codes, table, stream", "block_codec.cuh", r"^MYB_HD void huff_fast_emit_n", r"^MYB_HD void huff_fast_emit\("),
        ("tile_coord / block_row_col / copy_smem_to_global", "kernels.cu", r"^struct TileCoord", r"^MYB_D void copy_global_to_global\("),
    ] + _phases(ENC, ["ticket", "tile set-up, pixel loads (DCT itself is above)", "block sort", "hash clear, histogram call, statistics",
                      "queue", "plan dispatch", "size scan, reservation", "emit dispatch", "barrier", "copy out"]),
    "dct_decompress_kernel": [
        ("packed f32x2 helpers: the IDCT's products and sums, rounding constants", "kernels.cu", r"^struct __align__\(8\) f2", r"^MYB_D constexpr float dct_c"),
        ("idct_block / idct_block_tri: dequantise, transform, clamp, pack pixels", "kernels.cu", r"^MYB_D void idct_block\(", r"^struct SmemBytes"),
        ("SmemBytes / load_window (chunk bytes through shared addresses)", "kernels.cu", r"^struct SmemBytes", r"^struct SmemStream"),
        ("SmemStream::parse_table (PTX loop per table symbol)", "kernels.cu", r"^struct SmemStream", r"^  MYB_D void run\(DecStream"),
        ("SmemStream::run (PTX loop per stream symbol)", "kernels.cu", r"^  MYB_D void run\(DecStream", r"dec_tile_totals_kernel\("),
        ("huff_decode_fast + helpers (block_codec.cuh): header checks, code ranges, generic stream", "block_codec.cuh", r"^struct DecScratch", r"\Z"),
        ("cta_exclusive_scan / tile_coord", "kernels.cu", r"^struct TileCoord", r"^MYB_D void copy_smem_to_global"),
    ] + _phases(DEC, ["ticket", "descriptor, sizes, staging, scan, zero fill", "block sort", "entropy decoder call", "IDCT call, stores"]),
}


def resolve(regs):
    out = []
    for name, f, a, b, *after in regs:  # after: search both anchors behind the first line that matches it (the kernel's head)
        lines = (CSRC / f).read_text().split("\n")
        k0 = next((i for i, l in enumerate(lines) if re.search(after[0], l)), 0) if after else 0
        lo = next((i for i in range(k0, len(lines)) if re.search(a, lines[i])), None)
        if lo is None:
            continue
        hi = next((i for i in range(lo + 1, len(lines)) if re.search(b, lines[i])), len(lines)) if b != r"\Z" else len(lines)
        out.append((name, f, lo + 1, hi))
    return out


def profile_rows(rep, kernel):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    ks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(ks, ks[1:]):
        if kernel in rows[a][1]:
            hdr = rows[a + 1]
            ci, cs, ct, cx = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
            return rows[a][1], [(r[cx].strip(), int(r[ci] or 0), int(r[cs] or 0), int(r[ct] or 0)) for r in rows[a + 2:b] if len(r) > ct]
    raise SystemExit(f"no kernel {kernel} in {rep}")


def line_table(kernel_full_name):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", str(LIB)], cwd=tmp, capture_output=True)
        cubin = next(p for p in pathlib.Path(tmp).iterdir() if p.name.startswith("kernels") and "capi" not in p.name)
        text = subprocess.run(["nvdisasm", "--print-line-info", "-c", str(cubin)], capture_output=True, text=True).stdout
    want = "ILb0E" if "(bool)0" in kernel_full_name else "ILb1E" if "(bool)1" in kernel_full_name else ""
    base = re.search(r"(\w+_kernel)", kernel_full_name).group(1)
    lines = text.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and base in l and want in l)
    out, cur = [], ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.lstrip().startswith(".section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    name, prof = profile_rows(rep, kernel)
    dis = line_table(name)
    op = lambda s: [t for t in s.split() if not t.startswith("@")][0]
    if len(prof) != len(dis) or any(op(p[0]) != op(d[1]) for p, d in zip(prof, dis)):
        raise SystemExit(f"report ({len(prof)} instructions) and library ({len(dis)}) are not the same code")
    tot, tots = sum(p[1] for p in prof), sum(p[2] for p in prof)
    print(f"{name}: {len(prof)} instructions, {tot} warp instructions executed, {tots} stall samples")
    by_line = collections.defaultdict(lambda: [0, 0, 0, 0])
    for p, d in zip(prof, dis):
        v = by_line[d[0]]
        v[0] += p[1]; v[1] += p[2]; v[2] += 1; v[3] += p[3]
    regs = resolve(REGIONS.get(kernel.split("<")[0], []))
    agg = collections.OrderedDict((r[0], [0, 0, 0]) for r in regs)
    agg["everything else (intrinsics headers, atomics, shuffles)"] = [0, 0, 0]
    for (f, ln), v in by_line.items():
        for rname, rf, lo, hi in regs:
            if f == rf and lo <= ln <= hi:
                a = agg[rname]
                break
        else:
            a = agg["everything else (intrinsics headers, atomics, shuffles)"]
        a[0] += v[0]; a[1] += v[1]; a[2] += v[3]
    print("\nshare of executed warp instructions | share of stall samples | active lanes | region")
    for rname, v in agg.items():
        print(f"  {100 * v[0] / tot:5.1f} %  {100 * v[1] / max(tots, 1):5.1f} %  {v[2] / max(v[0], 1):4.1f}  {rname}")
    src = {f: (CSRC / f).read_text().split("\n") for f in ("kernels.cu", "block_codec.cuh") if (CSRC / f).exists()}
    print(f"\ntop {top} source lines (text from the CURRENT sources: shifted if they changed since the capture)")
    for (f, ln), v in sorted(by_line.items(), key=lambda x: -x[1][0])[:top]:
        text = src[f][ln - 1].strip()[:80] if f in src and ln <= len(src[f]) else ""
        print(f"  {100 * v[0] / tot:5.2f} %  {100 * v[1] / max(tots, 1):5.2f} %  {v[2]:4d} SASS  {v[3] / max(v[0], 1):4.1f} lanes  {f}:{ln}  {text}")


if __name__ == "__main__":
    main()
