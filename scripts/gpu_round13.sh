#!/bin/bash
# fused row-sum experiment (DGS_ATT_FUSEL=1: P V MMA with N = 80 = [O | L]): parity and kernel time
mkdir -p gpurun_out
DGS_ATT_FUSEL=1 timeout 100 python -m pytest tests/test_dit_gpu.py -q -x -s -k "attention" > gpurun_out/pytest_att_fusel.log 2>&1; echo "pytest(fusel) exit $?"
grep -E "passed|failed|FAILED|Error|rel=" gpurun_out/pytest_att_fusel.log | tail -4
for m in 1 0; do DGS_ATT_FUSEL=$m timeout 60 python tests/perf_kernels.py --attn-bwd 2>&1 | grep '"attention"' | head -1 | sed "s/^/fusel=$m /"; done
