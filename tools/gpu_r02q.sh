#!/bin/bash
# round 2, call q: skinny GEMM kernel (M <= 96): op tests, forward parity at small batches, batch sweep with / without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r02q_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -8 gpurun_out/r02q_pytest_ops.log | cut -c1-300
timeout 1200 python -m pytest tests/test_forward_gpu.py -m gpu -q -x > gpurun_out/r02q_pytest_fwd.log 2>&1; echo "pytest fwd rc=$?"; tail -8 gpurun_out/r02q_pytest_fwd.log | cut -c1-300
for sk in 1 0; do
  AFFT_GEMM_SKINNY=$sk timeout 300 python tools/batch_sweep.py ek100_sa_tsn fp16 1,2,4,5,8,32 > gpurun_out/r02q_sweep_skinny$sk.txt 2>&1; echo "sweep skinny=$sk rc=$?"; tail -16 gpurun_out/r02q_sweep_skinny$sk.txt | cut -c1-200
done
timeout 600 compute-sanitizer --tool memcheck python tools/ncu_forward.py 1 ek100_sa_tsn fp16 > gpurun_out/r02q_sanitize_b1.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r02q_sanitize_b1.log | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02q_launches_b1.csv python tools/ncu_forward.py 1 ek100_sa_tsn fp16 > gpurun_out/r02q_ncu_b1.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py launches gpurun_out/r02q_launches_b1.csv > gpurun_out/r02q_launches_b1_summary.txt 2>&1; head -24 gpurun_out/r02q_launches_b1_summary.txt | cut -c1-150
