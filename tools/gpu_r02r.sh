#!/bin/bash
# round 2, call r: skinny GEMM iteration - op tests, small-batch sweep for two row limits, launch list at B = 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm" > gpurun_out/r02r_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -4 gpurun_out/r02r_pytest_ops.log | cut -c1-300
for m in 96 32; do
  AFFT_GEMM_SKINNY_M=$m timeout 300 python tools/batch_sweep.py ek100_sa_tsn fp16 1,2,4,5 > gpurun_out/r02r_sweep_m$m.txt 2>&1; echo "sweep max_m=$m rc=$?"; grep '"max_ksplit": 4' gpurun_out/r02r_sweep_m$m.txt | cut -c1-200
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02r_launches_b1.csv python tools/ncu_forward.py 1 ek100_sa_tsn fp16 > gpurun_out/r02r_ncu_b1.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py launches gpurun_out/r02r_launches_b1.csv > gpurun_out/r02r_launches_b1_summary.txt 2>&1; head -16 gpurun_out/r02r_launches_b1_summary.txt | cut -c1-150
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02r_launches_b4.csv python tools/ncu_forward.py 4 ek100_sa_tsn fp16 > gpurun_out/r02r_ncu_b4.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py launches gpurun_out/r02r_launches_b4.csv > gpurun_out/r02r_launches_b4_summary.txt 2>&1; head -16 gpurun_out/r02r_launches_b4_summary.txt | cut -c1-150
