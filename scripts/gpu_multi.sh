#!/bin/bash
# multi-GPU evidence on N GPUs of one box (gpurun --gpus N): the multi-GPU tests (TESTS=1), then the strong-scaling bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8 | sort | uniq -c
if [ "${TESTS:-1}" = "1" ]; then
  timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short --timeout 1400 -p no:cacheprovider > gpurun_out/pytest_multi_${N}gpu.log 2>&1; tail -3 gpurun_out/pytest_multi_${N}gpu.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench exit $?"
grep "^{" gpurun_out/bench_${N}gpu.log > gpurun_out/r02_bench_${N}gpu.json
python - <<PY
import json
for l in open("gpurun_out/r02_bench_${N}gpu.json"):
    d = json.loads(l)
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_ceiling"], "per_rank", d["per_rank"])
    print("c4", {k: v for k, v in d.get("c4_loop_closure", {}).items() if k not in ("workload", "top10_sc_distance")})
PY
