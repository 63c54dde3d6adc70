#!/bin/bash
# last check of the round: every GPU test, smoke, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|exit|FAILED" gpurun_out/pytest_gpu.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; grep smoke gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','steps','warmup')}); print(d['roofline']['frac'], d['cpu_baseline']['value'])"
