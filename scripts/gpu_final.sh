#!/bin/bash
# round-end validation: every GPU test, smoke, the bench line (with CPU baseline), launch list, training step, other configs, C5 sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|exit|FAILED" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; grep smoke gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']); print(d['cpu_baseline']['value'], d['breakdown_ms']['families'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list exit $?"
timeout 300 python bench.py --workload train --batch 4 --steps 6 --warmup 3 > gpurun_out/bench_train_b4.json 2> gpurun_out/bench_train_b4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_b4.json')); print('train', d['value'], d['ms_per_step'], d['e2e'], d['breakdown_ms']['families'])"
timeout 300 python scripts/perf_configs.py gpurun_out/perf_configs.json --train > gpurun_out/perf_configs.log 2>&1; echo "configs exit $?"
timeout 200 python scripts/raster_sweep.py gpurun_out/raster_sweep.json > gpurun_out/raster_sweep.log 2>&1; echo "sweep exit $?"; tail -1 gpurun_out/raster_sweep.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:attention_fwd_kernel -s 6 -c 1 -f -o gpurun_out/prof_attention_fwd_kernel_v3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_att.log 2>&1; echo "ncu attention exit $?"
f=gpurun_out/prof_attention_fwd_kernel_v3.ncu-rep
ncu -i $f --page details 2>/dev/null | grep -v "^\s*$" | head -260 > gpurun_out/prof_attention_fwd_kernel_v3.details.txt
ncu -i $f --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed 2>/dev/null > gpurun_out/prof_attention_fwd_kernel_v3.metrics.csv
tail -2 gpurun_out/prof_attention_fwd_kernel_v3.metrics.csv | cut -c1-500
