#!/bin/bash
# CTA-pair GEMM as the default path: split-K weight gradients, fp32/aux TMA epilogues -> parity, training step, ncu capture
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dit_gpu.py tests/test_dit_bwd_gpu.py -q -x -s > gpurun_out/pytest_dit_default.log 2>&1; echo "pytest dit exit $?"
grep -E "passed|failed|FAILED|Error|gemm_tn|whole-gradient" gpurun_out/pytest_dit_default.log | cut -c1-220 | tail -14
timeout 300 python bench.py --workload train --batch 4 --steps 4 --warmup 3 > gpurun_out/bench_train_b4_v5.json 2> gpurun_out/bench_train_b4_v5.err
DGS_GEMM_SPLITK=0 timeout 300 python bench.py --workload train --batch 4 --steps 4 --warmup 3 > gpurun_out/bench_train_b4_nosplitk.json 2> gpurun_out/bench_train_b4_nosplitk.err
python - <<P
import json
for f in ("bench_train_b4_v5", "bench_train_b4_nosplitk"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 2), "samples/s", round(d["ms_per_step"], 2), "ms", {k: v for k, v in d["breakdown_ms"]["families"].items() if "gemm" in k}, "loss", d["loss"])
    except Exception as e:
        print(f, "failed", e)
P
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_bf16_2cta_kernel -s 8 -c 4 -f -o gpurun_out/prof_gemm_bf16_2cta_kernel_v2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_gemm2cta.log 2>&1; echo "ncu exit $?"
f=gpurun_out/prof_gemm_bf16_2cta_kernel_v2.ncu-rep
ncu -i $f --page details 2>/dev/null | grep -v "^\s*$" | head -400 > gpurun_out/prof_gemm_bf16_2cta_kernel_v2.details.txt
ncu -i $f --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum 2>/dev/null > gpurun_out/prof_gemm_bf16_2cta_kernel_v2.metrics.csv
cat gpurun_out/prof_gemm_bf16_2cta_kernel_v2.metrics.csv | cut -c1-400 | tail -6
