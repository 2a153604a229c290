#!/bin/bash
# round 2, call s: skinny GEMM (FIFO version + size rule + early weight prefetch): tests, sweep, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm" > gpurun_out/r02s_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -4 gpurun_out/r02s_pytest_ops.log | cut -c1-300
timeout 1200 python -m pytest tests/test_forward_gpu.py -m gpu -q > gpurun_out/r02s_pytest_fwd.log 2>&1; echo "pytest fwd rc=$?"; tail -12 gpurun_out/r02s_pytest_fwd.log | cut -c1-300
timeout 300 python tools/batch_sweep.py ek100_sa_tsn fp16 1,2,4,5,8 > gpurun_out/r02s_sweep.txt 2>&1; echo "sweep rc=$?"; grep '"max_ksplit": 4' gpurun_out/r02s_sweep.txt | cut -c1-200
AFFT_PDL=0 timeout 300 python tools/batch_sweep.py ek100_sa_tsn fp16 1 > gpurun_out/r02s_sweep_nopdl.txt 2>&1; grep '"max_ksplit": 4' gpurun_out/r02s_sweep_nopdl.txt | cut -c1-200
timeout 300 python tools/batch_sweep.py ek100_sa_tsn bf16 1 > gpurun_out/r02s_sweep_bf16.txt 2>&1; grep '"max_ksplit": 4' gpurun_out/r02s_sweep_bf16.txt | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s_launches_b1.csv python tools/ncu_forward.py 1 ek100_sa_tsn fp16 > gpurun_out/r02s_ncu_b1.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py launches gpurun_out/r02s_launches_b1.csv > gpurun_out/r02s_launches_b1_summary.txt 2>&1; head -16 gpurun_out/r02s_launches_b1_summary.txt | cut -c1-150
timeout 600 compute-sanitizer --tool memcheck python tools/ncu_forward.py 1 ek100_sa_tsn fp16 > gpurun_out/r02s_sanitize_b1.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r02s_sanitize_b1.log | tail -1
timeout 600 compute-sanitizer --tool racecheck python tools/ncu_forward.py 1 ek100_sa_tsn fp16 > gpurun_out/r02s_racecheck_b1.log 2>&1; grep -E "RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/r02s_racecheck_b1.log | tail -2
