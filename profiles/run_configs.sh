#!/bin/bash
# One 1-GPU gpurun call: BASELINE configs[0..2] with their parity gates (configs_final.py) and an `ncu --set full` capture of the kernels
# beside the two codec tile kernels (other_kernels_pass.py).  Outputs in gpurun_out/; what is kept goes to profiles/.
set -u
O=gpurun_out
timeout 300 python profiles/configs_final.py > $O/r02_configs.json 2> $O/configs.err; echo "configs rc=$?"; tail -c 400 $O/configs.err
timeout 100 python profiles/other_kernels_pass.py > $O/other_pass.log 2>&1; rc=$?; echo "pass rc=$rc"; tail -c 300 $O/other_pass.log
if [ $rc -eq 0 ]; then
  timeout 240 ncu --set full --clock-control none --profile-from-start off -c 40 \
    -k regex:'xrgb_to_iyuv_kernel|bgr24_to_iyuv_kernel|iyuv_to_rgba_kernel|heavy|place_|finalize_' \
    -o $O/r02_other -f python profiles/other_kernels_pass.py > $O/ncu_other.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu_other.log
  python profiles/summarize_ncu.py $O/r02_other.ncu-rep > $O/r02_ncu_other_kernels.json; grep -c kernel $O/r02_ncu_other_kernels.json
fi
