mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo rc=$?
python -c "import json;d=json.load(open('gpurun_out/bench_8gpu.json'));print(d['value'],d['e2e']['value'],d['n_gpus'])"
