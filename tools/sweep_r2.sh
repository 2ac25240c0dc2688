#!/bin/bash
# BASELINE.json configs[4]: num_envs sweep on one B200 (CoordSum 3x10-30 and the LBF bench env), round-2 build. -> gpurun_out/sweep_r2.txt
mkdir -p gpurun_out; : > gpurun_out/sweep_r2.txt
for env in coordsum lbf; do
  for e in 1024 4096 16384 65536; do
    if [ $e = 65536 ] && [ $env = lbf ]; then continue; fi
    timeout 400 python bench.py --env $env --num-envs $e --steps 3 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/sw.json 2>/dev/null
    python - <<PY >> gpurun_out/sweep_r2.txt
import json
try:
    d=json.loads(open("gpurun_out/sw.json").read().strip().splitlines()[-1]); print("$env", $e, round(d["ms_per_step"],1), round(d["value"]/1e6,2), d["gpu_launches"])
except Exception as ex: print("$env", $e, "failed", ex)
PY
  done
done
cat gpurun_out/sweep_r2.txt
