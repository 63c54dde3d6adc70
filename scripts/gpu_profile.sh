#!/bin/bash
# ncu --set full capture of the hot kernels (one GPU, short commands) -> gpurun_out/prof_*.ncu-rep + text summaries.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh'
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# inference step (bench.py): forward attention, the default GEMM, the blend kernel, the instance sort
for pat in attention_fwd_kernel gemm_bf16_kernel blend_forward_kernel DeviceRadixSortOnesweepKernel ln_modulate_kernel; do
  timeout 300 $NCU -k regex:$pat -s 6 -c 2 -f -o gpurun_out/prof_$pat \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$pat.log 2>&1
  echo "ncu $pat exit $?"
done
# training kernels: attention backward pair (micro-benchmark), blend backward + transposes (2-layer training step)
for pat in attn_bwd_dq_kernel attn_bwd_dkv_kernel; do
  timeout 300 $NCU -k regex:$pat -s 3 -c 1 -f -o gpurun_out/prof_$pat python tests/perf_kernels.py --attn-bwd > gpurun_out/prof_$pat.log 2>&1
  echo "ncu $pat exit $?"
done
for pat in blend_backward_kernel transpose_kernel; do
  DGS_LAYERS=2 timeout 300 $NCU -k regex:$pat -s 2 -c 1 -f -o gpurun_out/prof_$pat python tests/perf_train.py gpurun_out/x.json 1 10 1 > gpurun_out/prof_$pat.log 2>&1
  echo "ncu $pat exit $?"
done
for f in gpurun_out/prof_*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page details 2>/dev/null | grep -v "^\s*$" | head -220 > gpurun_out/$b.details.txt
  ncu -i $f --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed 2>/dev/null > gpurun_out/$b.metrics.csv
done
ls -la gpurun_out/*.ncu-rep
