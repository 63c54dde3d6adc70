#!/bin/bash
# last 1-GPU check of the round: the raster tests (new view-chunk test), bench line with the pinned result read-back
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_raster_gpu.py -q -x -s -k "chunked or batched or c1_forward" > gpurun_out/pytest_raster_chunk.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|chunked d" gpurun_out/pytest_raster_chunk.log | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']['frac'], d['cpu_baseline']['value'])"
