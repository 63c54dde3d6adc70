#!/bin/bash
# One parameterised GPU-box script (replaces the per-call gpu_roundN.sh files of round 1).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_call.sh <steps...>'
# steps: tests [pytest-args]   every GPU test (or the given selection)
#        smoke                 __graft_entry__.smoke()
#        bench [N]             bench.py on N GPUs (default 1; N > 1 through torchrun), JSON -> gpurun_out/bench_nN.json
#        refarm                bench.py --impl reference
#        launches              ncu launch list of the bench step
#        train B [N] [flags]   bench.py --workload train --batch B on N GPUs
#        kernels               tests/perf_kernels.py --all
#        ncu PATTERN           ncu --set full capture of one kernel of the bench step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29577"
while [ $# -gt 0 ]; do
  step=$1; shift
  case $step in
    tests)
      sel="tests"; if [ $# -gt 0 ] && [[ "$1" == tests/* ]]; then sel=$1; shift; fi
      timeout 1200 python -m pytest $sel -q -m gpu -s --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
      grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | head -40 ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; grep smoke gpurun_out/smoke.log ;;
    bench)
      n=1; if [[ "$1" =~ ^[0-9]+$ ]]; then n=$1; shift; fi
      if [ "$n" = 1 ]; then timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
      else NCCL_DEBUG=INFO timeout 900 $TR --nproc-per-node $n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.out 2> gpurun_out/bench_n$n.err
           grep '^{"metric' gpurun_out/bench_n$n.out | tail -1 > gpurun_out/bench_n$n.json; grep -c "nranks $n" gpurun_out/bench_n$n.out gpurun_out/bench_n$n.err; fi
      echo "bench n=$n exit $?"; python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$n.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks")}); print(d.get("roofline")); print(d["breakdown_ms"]["families"])
print("cpu", d.get("cpu_baseline") and d["cpu_baseline"]["value"]); print("train", d.get("train")); print("per_rank", d.get("per_rank"))
PY
      tail -3 gpurun_out/bench_n$n.err ;;
    refarm)
      timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list exit $?" ;;
    train)
      b=$1; shift; n=1; if [[ "$1" =~ ^[0-9]+$ ]]; then n=$1; shift; fi
      fl=""; while [ $# -gt 0 ] && [[ "$1" == --* ]]; do fl="$fl $1"; shift; done
      tag="b${b}_n${n}$(echo $fl | tr -d ' -')"
      if [ "$n" = 1 ]; then timeout 600 python bench.py --workload train --batch $b --steps 6 --warmup 3 $fl > gpurun_out/train_$tag.json 2> gpurun_out/train_$tag.err
      else timeout 600 $TR --nproc-per-node $n bench.py --gpus $n --workload train --batch $b --steps 6 --warmup 3 $fl > gpurun_out/train_$tag.out 2> gpurun_out/train_$tag.err
           grep '^{"metric' gpurun_out/train_$tag.out | tail -1 > gpurun_out/train_$tag.json; fi
      python - <<PY
import json
d = json.load(open("gpurun_out/train_$tag.json"))
print("train $tag", round(d["value"], 2), "samples/s", round(d["ms_per_step"], 2), "ms", "exposed", d.get("allreduce_exposed_ms"), "mem", round(d["mem_gb"], 1), d["breakdown_ms"]["families"])
PY
      tail -2 gpurun_out/train_$tag.err ;;
    kernels)
      timeout 200 python tests/perf_kernels.py --all 2>&1 | tee gpurun_out/perf_kernels.txt ;;
    ncu)
      pat=$1; shift
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s 6 -c 1 -f -o gpurun_out/prof_$pat python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$pat.log 2>&1; echo "ncu $pat exit $?"
      ncu -i gpurun_out/prof_$pat.ncu-rep --page details 2>/dev/null | grep -v "^\s*$" | head -260 > gpurun_out/prof_$pat.details.txt
      ncu -i gpurun_out/prof_$pat.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed 2>/dev/null > gpurun_out/prof_$pat.metrics.csv
      tail -2 gpurun_out/prof_$pat.metrics.csv | cut -c1-400 ;;
    *) echo "unknown step $step" ;;
  esac
done
