#!/bin/bash
# One 8-GPU gpurun call: sharded-path test log, shard8k bench at N = 1, 2, 4, 8, host copy ceiling at N = 1, 2, 4, 8.
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_shard_device.py -x -q -m gpu 2>&1 | tail -4
N=$(nvidia-smi -L | wc -l)
cp profiles/r02_shard8k_test_${N}gpu.log $O/ 2>/dev/null
timeout 200 python bench.py --workload shard8k --steps 20 > $O/r02_shard8k_1gpu.json 2> $O/shard1.err; echo "shard N=1 rc=$?"
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 300 $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --workload shard8k --gpus $n --steps 20 > $O/r02_shard8k_${n}gpu.json 2> $O/shard$n.err; echo "shard N=$n rc=$?"
done
if [ 2 -le $N ]; then timeout 600 $TR --nproc-per-node 2 --master-port 29555 bench.py --gpus 2 > $O/r02_bench_2gpu.json 2> $O/bench2gpu.err; echo "bench N=2 rc=$?"; fi
: > $O/r02_pcie_ceiling.jsonl
timeout 120 python profiles/pcie_ceiling.py >> $O/r02_pcie_ceiling.jsonl 2>$O/pcie.err
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 200 $TR --nproc-per-node $n --master-port $((29700+n)) profiles/pcie_ceiling.py >> $O/r02_pcie_ceiling.jsonl 2>>$O/pcie.err
done
cat $O/r02_pcie_ceiling.jsonl
for n in 1 2 4 8; do python - <<PY
import json
try:
    l=json.load(open("$O/r02_shard8k_${n}gpu.json")); print($n, l["value"], l["shard"]["compress_us_per_image"], l["shard"]["decompress_us_per_image"], l["shard"]["exchange_plus_assemble_us"], l["e2e"]["ms_per_image"], l["parity_checked"])
except Exception as e: print($n, "no line", e)
PY
done
