#!/bin/bash
# Line coverage of csrc/block_codec.cuh (the entropy coder / decoder source the kernels inline) by the host tests:
# tests/hostemu/hostemu.cpp built with --coverage, tests/test_hostemu.py run against it, gcov on the header, lines that no
# instantiation ever executed listed.  CPU only.   usage: bash profiles/hostemu_coverage.sh   (result: profiles/r02_hostemu_coverage.txt)
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=$(mktemp -d)
/usr/bin/g++ -std=gnu++17 -O0 -g --coverage -fPIC -shared -ffp-contract=off -frounding-math -o $T/libhostemu_cov.so $ROOT/tests/hostemu/hostemu.cpp
(cd $ROOT && HOSTEMU_LIB=$T/libhostemu_cov.so python -m pytest tests/test_hostemu.py -q -p no:cacheprovider | tail -1)
(cd $T && gcov -o $T libhostemu_cov.so-hostemu.gcno > gcov.log 2>/dev/null; grep -A1 "block_codec.cuh" gcov.log | head -2)
python3 - $T/block_codec.cuh.gcov <<'PY'
import re, sys
best, src = {}, {}
for l in open(sys.argv[1], errors="replace"):
    m = re.match(r"\s*([#=\-\d\*]+):\s*(\d+):(.*)", l)
    if not m or m.group(1) == "-" or m.group(2) == "0":
        continue
    ln = int(m.group(2))
    c = 0 if m.group(1)[0] in "#=" else int(m.group(1).rstrip("*"))
    best[ln] = max(best.get(ln, 0), c)
    src.setdefault(ln, m.group(3).strip())
never = [ln for ln in sorted(best) if best[ln] == 0]
print(f"{len(best)} source lines with code, {len(best) - len(never)} executed by at least one instantiation, {len(never)} never:")
for ln in never:
    print(f"  block_codec.cuh:{ln}  {src[ln][:110]}")
PY
rm -rf $T
