#!/bin/bash
# 2-GPU box: column-sum kernel parity + training step on 1 GPU, then the 2-GPU lines (replicas / NCCL gradient all-reduce)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dit_bwd_gpu.py -q -x -k "dit_backward or optimizer or train" > gpurun_out/pytest_bwd_colsum.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED" gpurun_out/pytest_bwd_colsum.log | tail -3
timeout 300 python bench.py --workload train --batch 4 --steps 6 --warmup 3 > gpurun_out/bench_train_b4_v7.json 2> gpurun_out/bench_train_b4_v7.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_v3.json 2> gpurun_out/bench_2gpu_v3.err
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --workload train --batch 4 --steps 6 --warmup 3 > gpurun_out/bench_train_2gpu_b4_v2.json 2> gpurun_out/bench_train_2gpu_b4_v2.err
python - <<P
import json
for f in ("bench_train_b4_v7", "bench_2gpu_v3", "bench_train_2gpu_b4_v2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms e2e", round(d["e2e"]["value"], 2), {k: v for k, v in d["breakdown_ms"]["families"].items() if "elementwise" in k})
    except Exception as e:
        print(f, "failed", e)
P
tail -3 gpurun_out/bench_2gpu_v3.err gpurun_out/bench_train_2gpu_b4_v2.err | cut -c1-300
