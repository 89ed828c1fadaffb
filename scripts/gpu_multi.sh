#!/bin/bash
# multi-GPU checks: sharded loop-closure test + the strong-scaling bench on N GPUs (N = $1)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short --timeout 1400 -p no:cacheprovider -x  > gpurun_out/pytest_multi.log 2>&1; tail -5 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --pairs ${PAIRS:-4096} --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench exit $?"
grep -v "^{" gpurun_out/bench_${N}gpu.log | tail -5
python - <<PY
import json
for l in open("gpurun_out/bench_${N}gpu.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "per_rank", d["per_rank"], "scaling", d["scaling"])
        print("c4", {k: v for k, v in d.get("c4_loop_closure", {}).items() if k != "workload"})
PY
