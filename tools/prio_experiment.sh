for m in 0 1 2; do
  MAGPO_PRIO_MODE=$m timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_prio${m}_lbf.json 2>/dev/null
  MAGPO_PRIO_MODE=$m timeout 200 python bench.py --env rware --num-envs 1024 --update-batch-size 1 --steps 4 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_prio${m}_rware.json 2>/dev/null
  python - <<PY
import json
for e in ("lbf","rware"):
    d=json.load(open("gpurun_out/bench_prio${m}_%s.json"%e)); print("prio_mode=${m}", e, round(d["ms_per_step"],2), d["phase_ms"])
PY
done
