for r in 128 64 32; do
  MAGPO_GRU_RPT=$r timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rpt${r}_lbf.json 2>/dev/null
  MAGPO_GRU_RPT=$r timeout 200 python bench.py --env rware --num-envs 1024 --update-batch-size 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rpt${r}_rware.json 2>/dev/null
  python - <<PY
import json
for e in ("lbf","rware"):
    d=json.load(open("gpurun_out/bench_rpt${r}_%s.json"%e)); print("rpt=${r}", e, round(d["ms_per_step"],1), d["breakdown_ms_per_step"]["gru_pointwise"]["ms"])
PY
done
MAGPO_GRU_RPT=32 timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_networks.py tests/test_gpu_lbf.py -q -m gpu -x 2>&1 | tail -2
MAGPO_GRU_RPT=64 timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_networks.py -q -m gpu -x 2>&1 | tail -2
