#!/bin/bash
# FIRST call of the next round: attention timeline probe (extended: per-MMA-group issue times) at the bench shape, the
# training shape and the 512^2 shape; then the standard check.  ~1 min of GPU time.
mkdir -p gpurun_out
timeout 30 scripts/bin/att_probe 4098 16 1 > gpurun_out/att_probe_v2.txt 2>&1
timeout 30 scripts/bin/att_probe 4098 16 4 >> gpurun_out/att_probe_v2.txt 2>&1
timeout 60 scripts/bin/att_probe 16386 16 1 >> gpurun_out/att_probe_v2.txt 2>&1
cat gpurun_out/att_probe_v2.txt
