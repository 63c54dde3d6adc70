#!/bin/bash
# near-fraction of the two-phase binning (1/8 default): bench step and the dense half of the C5 sweep at 1/8, 1/16, 1/32
mkdir -p gpurun_out
for k in 3 4 5; do
  DGS_RASTER_NEAR_LOG2=$k timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_near$k.json 2> gpurun_out/bench_near$k.err
  DGS_RASTER_NEAR_LOG2=$k timeout 120 python scripts/raster_sweep.py gpurun_out/sweep_near$k.json --quick > gpurun_out/sweep_near$k.log 2>&1
  python - <<P
import json
d = json.load(open("gpurun_out/bench_near$k.json"))
f = d["breakdown_ms"]["families"]
print("near_log2=$k", round(d["value"], 2), "steps/s raster", {k: v for k, v in f.items() if k.startswith("raster")})
s = json.load(open("gpurun_out/sweep_near$k.json"))
print("   sweep", [(c["dist"], c["P"], c["res"], round(c["fwd_ms"], 2), round(c["fwdbwd_ms"], 2)) for c in s["cases"]])
P
done
