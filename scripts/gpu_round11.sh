#!/bin/bash
# stale-max softmax experiment (DGS_ATT_STALE=1): parity and kernel time vs the default
mkdir -p gpurun_out
DGS_ATT_STALE=1 timeout 300 python -m pytest tests/test_dit_gpu.py tests/test_dit_bwd_gpu.py -q -x -s -k "attention" > gpurun_out/pytest_att_stale.log 2>&1; echo "pytest(stale) exit $?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_att_stale.log | tail -3
for m in 0 1 0 1; do DGS_ATT_STALE=$m timeout 120 python tests/perf_kernels.py --attn-bwd 2>&1 | grep '"attention"' | head -1 | sed "s/^/stale=$m /"; done
DGS_ATT_STALE=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stale1.json 2> gpurun_out/bench_stale1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_stale1.json')); print('bench stale=1', round(d['value'],2), 'att', d['breakdown_ms']['families']['dit.attention'])"
