#!/bin/bash
# two-softmax-threads-per-row attention forward (DGS_ATT_TPR2=1): parity and kernel time vs the default
mkdir -p gpurun_out
DGS_ATT_TPR2=1 timeout 200 python -m pytest tests/test_dit_gpu.py tests/test_dit_bwd_gpu.py -q -x -s -k "attention" > gpurun_out/pytest_att_tpr2.log 2>&1; echo "pytest(tpr2) exit $?"
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/pytest_att_tpr2.log | tail -4
for m in 0 1; do DGS_ATT_TPR2=$m timeout 100 python tests/perf_kernels.py --attn-bwd 2>&1 | grep '"attention"' | head -1 | sed "s/^/tpr2=$m /"; done
DGS_ATT_TPR2=1 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tpr2.json 2> gpurun_out/bench_tpr2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tpr2.json')); print('bench tpr2=1', round(d['value'],2), 'att', d['breakdown_ms']['families']['dit.attention'])"
