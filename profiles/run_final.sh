#!/bin/bash
# One 1-GPU gpurun call that regenerates the evidence of the final code: test log, ncu --set full summaries of both codec kernels
# (synthetic and natural frames), the launch list of one bench step with the DRAM bytes of the codec kernels, the bench line
# (which then reports that traffic) and the reference arm.  Everything lands in gpurun_out/; copy what is kept to profiles/.
set -u
O=gpurun_out
K='regex:dct_compress_kernel|dct_decompress_kernel'
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/r02_pytest_final.log; cat $O/r02_pytest_final.log
for w in ng:50 nat:50; do
  n=${w%%:*}
  timeout 400 ncu --set full --import-source on --clock-control none -k "$K" -s 2 -c 2 -o $O/r02f_$n -f python profiles/one_pass.py $w > $O/ncu_$n.log 2>&1
  python profiles/summarize_ncu.py $O/r02f_$n.ncu-rep > $O/r02_ncu_${n}_8frames.json
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02_launches_raw.csv \
  -k regex:'dct_|heavy|scan_|place_|finalize_|parse_|dec_|publish_|sm_copy|iyuv' python bench.py --steps 1 --warmup 1 --no-e2e --no-sweep --no-cpu-baseline > /dev/null 2> $O/ncu_launches.err
python profiles/launch_list.py $O/r02_launches_raw.csv $O/r02_launches_64frames.csv $O/r02_traffic.json $O/r02_ncu_ng_8frames.json "8 synthetic frames" | tee $O/r02_launch_list.txt
cp $O/r02_traffic.json profiles/r02_traffic.json
ONE_PASS_FRAMES=16 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/raw_nat16.csv -k regex:'dct_|heavy|scan_|place_|finalize_|parse_|dec_|publish_' python profiles/one_pass.py nat:50 nat:90 > /dev/null 2>&1
ONE_PASS_FRAMES=16 python profiles/one_pass.py --parse $O/raw_nat16.csv nat:50 nat:90 | grep -v "array<\|_cuda\|CUDAFunctor\|launch_clamp\|arange" > $O/r02_launches_natural_16frames.txt
timeout 900 python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err; echo "bench rc=$?"; tail -c 600 $O/r02_bench_final.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "reference rc=$?"; cat $O/r02_bench_reference_arm.json | cut -c1-400
timeout 300 python bench.py --workload shard8k --steps 20 > $O/r02_shard8k_1gpu.json 2> $O/shard1.err; echo "shard rc=$?"
timeout 400 python profiles/phase_clocks.py > $O/r02_phase_clocks_final.json 2> $O/phase.err; echo "phase clocks rc=$?"
