#!/bin/bash
# ncu --set full capture of the hot kernels (one GPU, short command) -> gpurun_out/prof_*.ncu-rep
mkdir -p gpurun_out
for pat in attention_fwd_kernel gemm_bf16_2cta_kernel blend_forward_kernel DeviceRadixSortOnesweepKernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 6 -c 2 -f -o gpurun_out/prof_$pat \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$pat.log 2>&1
  echo "ncu $pat exit $?"
done
ls -la gpurun_out/*.ncu-rep
