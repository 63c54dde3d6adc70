#!/bin/bash
# FIRST call of the next round (~2 min of GPU time):
#  1. the elect.sync issue path: parity with DGS_ATT_UNI=1 / DGS_GEMM_UNI=1, kernel times with
#     and without, bench step with both on;
#  2. the attention timeline probe (per-role wait cycles, per-MMA-group issue times) at the bench / training / 512^2 shapes.
mkdir -p gpurun_out
DGS_ATT_UNI=1 DGS_GEMM_UNI=1 timeout 400 python -m pytest tests/test_dit_gpu.py tests/test_dit_bwd_gpu.py -q -x > gpurun_out/pytest_uni.log 2>&1; echo "pytest(uni) exit $?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_uni.log | tail -3
for m in 0 1; do
  DGS_ATT_UNI=$m DGS_GEMM_UNI=$m timeout 120 python tests/perf_kernels.py --all 2>&1 | sed "s/^/uni=$m /"
done
for m in 0 1; do
  DGS_ATT_UNI=$m DGS_GEMM_UNI=$m timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_uni$m.json 2> gpurun_out/bench_uni$m.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_uni$m.json')); print('bench uni=$m', round(d['value'],2), 'steps/s', {k: v for k, v in d['breakdown_ms']['families'].items() if 'gemm' in k or 'attention' in k})"
done
timeout 30 scripts/bin/att_probe 4098 16 1 > gpurun_out/att_probe_v2.txt 2>&1
timeout 30 scripts/bin/att_probe 4098 16 4 >> gpurun_out/att_probe_v2.txt 2>&1
timeout 60 scripts/bin/att_probe 16386 16 1 >> gpurun_out/att_probe_v2.txt 2>&1
cat gpurun_out/att_probe_v2.txt
timeout 300 python scripts/library_baseline.py gpurun_out/r2_library_baseline.json > gpurun_out/library_baseline.log 2>&1; tail -30 gpurun_out/library_baseline.log
