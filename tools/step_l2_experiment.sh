#!/bin/bash
# A/B of the decoder-state L2 policy in the rollout step kernel (MAGPO_STEP_KEEP_L2 = 0 / 1, csrc/sable_step.cu load_rows<KEEP>).
# Outputs under gpurun_out/.
mkdir -p gpurun_out
for k in 1 0 1 0; do
  MAGPO_STEP_KEEP_L2=$k timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l2_${k}.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_l2_${k}.json")); print("lbf mode=${k}", round(d["ms_per_step"],1), d["phase_ms"], "sample", d["breakdown_ms_per_step"]["sample"]["ms"])
PY
done
for k in 1 0 1 0; do
  MAGPO_STEP_KEEP_L2=$k timeout 200 python bench.py --env rware --num-envs 1024 --update-batch-size 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l2_${k}_rware.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_l2_${k}_rware.json")); print("rware mode=${k}", round(d["ms_per_step"],1), d["phase_ms"], "sample", d["breakdown_ms_per_step"]["sample"]["ms"])
PY
done
timeout 600 python -m pytest tests/test_gpu_system.py tests/test_gpu_rware.py tests/test_gpu_lbf.py tests/test_gpu_baseline_configs.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
