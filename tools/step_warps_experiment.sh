for w in 14 7 0; do
  MAGPO_STEP_WARPS=$w timeout 200 python bench.py --env rware --num-envs 1024 --update-batch-size 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sw${w}_rware.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_sw${w}_rware.json")); print("step_warps=${w}", round(d["ms_per_step"],1), d["phase_ms"], d["breakdown_ms_per_step"]["sample"]["ms"], d["breakdown_ms_per_step"]["gru_pointwise"]["ms"])
PY
done
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sw_lbf.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_sw_lbf.json')); print('lbf', round(d['ms_per_step'],1), d['phase_ms'])"
timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_rware.py tests/test_gpu_lbf.py -q -m gpu -x 2>&1 | tail -2
