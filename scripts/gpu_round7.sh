#!/bin/bash
# per-kernel launch list of a 2-layer training step at batch 4 (where do the backward glue milliseconds go?)
mkdir -p gpurun_out
DGS_LAYERS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_2layers_b4.csv python tests/perf_train.py gpurun_out/perf_train_2l.json 4 10 1 > gpurun_out/perf_train_2l.log 2>&1; echo "ncu exit $?"
tail -2 gpurun_out/perf_train_2l.log | cut -c1-300
