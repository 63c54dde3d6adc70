#!/bin/bash
# attention forward: key-axis split of the last partial wave -> parity, kernel time, step time
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dit_gpu.py tests/test_dit_bwd_gpu.py -q -x -s -k "attention or denoiser or dit_backward" > gpurun_out/pytest_att_split.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|FAILED|Error|attention B=|attention bwd" gpurun_out/pytest_att_split.log | cut -c1-200 | tail -24
for m in 0 1; do
  DGS_ATT_SPLIT=$m timeout 120 python tests/perf_kernels.py --attn-bwd 2>&1 | grep attention | head -2
done
for m in 0 1; do
  DGS_ATT_SPLIT=$m timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attsplit$m.json 2> gpurun_out/bench_attsplit$m.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_attsplit$m.json"))
    print("bench att_split=$m", round(d["value"], 2), "steps/s e2e", round(d["e2e"]["value"], 2), "att ms", d["breakdown_ms"]["families"]["dit.attention"], "roof", d["roofline"]["frac"])
except Exception as e:
    print("bench att_split=$m failed", e)
P
done
