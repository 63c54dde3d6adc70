#!/bin/bash
# one combined call: GEMM probe v2, CTA-pair GEMM parity + perf, attention at N=16386, raster sweep (C5), other configs
mkdir -p gpurun_out
timeout 120 scripts/bin/gemm_probe > gpurun_out/gemm_probe_v2.txt 2>&1; echo "probe exit $?"
DGS_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_dit_gpu.py -q -x -k "gemm or attention" -s > gpurun_out/pytest_2cta.log 2>&1; echo "pytest 2cta exit $?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_2cta.log | tail -5
timeout 200 python -m pytest tests/test_dit_bwd_gpu.py -q -x -k "attention_backward" -s > gpurun_out/pytest_attn_bwd.log 2>&1; echo "pytest attn bwd exit $?"
grep -E "passed|failed|FAILED|N=16386" gpurun_out/pytest_attn_bwd.log | tail -5
timeout 120 python tests/perf_kernels.py > gpurun_out/perf_kernels_1cta.txt 2>&1
DGS_GEMM_2CTA=1 timeout 120 python tests/perf_kernels.py > gpurun_out/perf_kernels_2cta.txt 2>&1
cat gpurun_out/perf_kernels_1cta.txt gpurun_out/perf_kernels_2cta.txt
timeout 280 python scripts/raster_sweep.py gpurun_out/raster_sweep.json > gpurun_out/raster_sweep.log 2>&1; echo "sweep exit $?"; tail -2 gpurun_out/raster_sweep.log
timeout 420 python scripts/perf_configs.py gpurun_out/perf_configs.json --train > gpurun_out/perf_configs.log 2>&1; echo "configs exit $?"; tail -3 gpurun_out/perf_configs.log | cut -c1-600
