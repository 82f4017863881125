#!/usr/bin/env python3
"""Static facts about every kernel of the built library, from cuobjdump (CPU only): SASS instruction count, registers, stack (spill)
bytes, static shared memory, the most frequent opcodes, and the mnemonics that identify what the code is made of -- packed FP32
(FMUL2 / FFMA2 / FADD2), scalar FFMA (must be 0 in the codec kernels: a contraction would break bit-exactness), the TMA bulk copy
(UBLKCP) and its mbarrier (SYNCS), warp collectives (REDUX, VOTE, SHFL, MATCH), atomics.
usage: python profiles/sass_stats.py > profiles/r02_sass_stats.json"""
import collections, json, pathlib, re, subprocess, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
LIB = ROOT / "yuv-manipulations-2_b200" / "lib" / "libmyyuvb200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
    usage[m.group(1)] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL)[: ]+(\d+)", m.group(2))}
out = {}
for part in sass.split("Function : ")[1:]:
    name = part.split("\n", 1)[0].strip()
    ops = collections.Counter()
    for line in part.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            ops[m.group(1)] += 1
    n = sum(ops.values())
    pick = lambda *names: {k: ops[k] for k in names if ops[k]}
    full = demangle(name).replace("void ", "").replace("myyuvb::", "")
    m = re.match(r"(\w+(?:<.*?>)?)\(", full)
    short = (m.group(1) if m else full).replace("(bool)0", "queue").replace("(bool)1", "in place").replace("(int)", "")
    out[short] = {"instructions": n, **{k.lower(): v for k, v in usage.get(name, {}).items()},
                  "packed_fp32": pick("FMUL2", "FFMA2", "FADD2"), "scalar_ffma": ops["FFMA"],
                  "tma_bulk_copy": pick("UBLKCP", "SYNCS"), "warp_collectives": pick("REDUX", "VOTE", "SHFL", "MATCH", "WARPSYNC"),
                  "atomics": pick("ATOMG", "ATOMS", "RED", "ATOM"), "barriers": ops["BAR"],
                  "shared_memory_ops": pick("LDS", "STS"), "global_memory_ops": pick("LDG", "STG"), "local_memory_ops": pick("LDL", "STL"),
                  "top_opcodes": dict(ops.most_common(8))}
json.dump(out, sys.stdout, indent=1)
