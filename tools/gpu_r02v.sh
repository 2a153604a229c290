#!/bin/bash
# round 2, call v (2 GPUs): the driver's scaling launch of the default bench (both arms) at N = 2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02v_bench_n2.json 2> gpurun_out/r02v_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02v_bench_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02v_ref_n2.json 2> gpurun_out/r02v_ref_n2.err; echo "reference n2 rc=$?"; cut -c1-300 gpurun_out/r02v_ref_n2.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02v_bench_n2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "scaling", "gpu_launches", "dtype")}, "e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("keys", sorted(d.keys()))
PY
