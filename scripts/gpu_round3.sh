#!/bin/bash
# CTA-pair GEMM (fixed protocol, MN-major wgrad, TMA epilogues): probe, parity, step-level effect; raster overflow path; configs
mkdir -p gpurun_out
timeout 120 scripts/bin/gemm_probe > gpurun_out/gemm_probe_v3.txt 2>&1; echo "probe exit $?"
grep -E "check|FAILED|error" gpurun_out/gemm_probe_v3.txt | head -20
DGS_GEMM_2CTA=1 timeout 400 python -m pytest tests/test_dit_gpu.py tests/test_dit_bwd_gpu.py -q -x -s > gpurun_out/pytest_2cta_full.log 2>&1; echo "pytest 2cta exit $?"
grep -E "passed|failed|FAILED|Error|rel=" gpurun_out/pytest_2cta_full.log | tail -8
timeout 300 python -m pytest tests/test_raster_gpu.py tests/test_diffusion.py -q -x -m gpu > gpurun_out/pytest_raster_diff.log 2>&1; echo "pytest raster+diffusion exit $?"
tail -3 gpurun_out/pytest_raster_diff.log
DGS_GEMM_2CTA=1 timeout 120 python tests/perf_kernels.py > gpurun_out/perf_kernels_2cta_tma.txt 2>&1
DGS_GEMM_2CTA=1 DGS_GEMM_TMA_EPI=0 timeout 120 python tests/perf_kernels.py > gpurun_out/perf_kernels_2cta_notma.txt 2>&1
cat gpurun_out/perf_kernels_2cta_tma.txt gpurun_out/perf_kernels_2cta_notma.txt | grep gemm
for m in 0 1; do
  DGS_GEMM_2CTA=$m timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2cta$m.json 2> gpurun_out/bench_2cta$m.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_2cta$m.json"))
    print("bench 2cta=$m", round(d["value"], 2), "steps/s e2e", round(d["e2e"]["value"], 2), {k: v for k, v in d["breakdown_ms"]["families"].items() if "gemm" in k or "attention" in k or "ln" in k})
except Exception as e:
    print("bench 2cta=$m failed", e)
P
done
DGS_GEMM_2CTA=1 timeout 300 python bench.py --workload train --batch 4 --steps 4 --warmup 3 > gpurun_out/bench_train_2cta1.json 2> gpurun_out/bench_train_2cta1.err
python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_train_2cta1.json"))
    print("train 2cta=1", round(d["value"], 2), "samples/s", d["ms_per_step"], {k: v for k, v in d["breakdown_ms"]["families"].items() if "gemm" in k})
except Exception as e:
    print("train bench failed", e)
P
timeout 420 python scripts/perf_configs.py gpurun_out/perf_configs.json --train > gpurun_out/perf_configs.log 2>&1; echo "configs exit $?"; tail -2 gpurun_out/perf_configs.log | cut -c1-400
