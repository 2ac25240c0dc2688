for e in 1024 4096 16384; do
  timeout 280 python bench.py --num-envs $e --steps 3 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_sweep_lbf_$e.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_sweep_lbf_$e.json")); print("lbf num_envs=$e", round(d["ms_per_step"],2), round(d["value"]), d["phase_ms"])
PY
done
timeout 200 python bench.py --env coordsum --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_coordsum.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_final_coordsum.json')); print('coordsum', round(d['ms_per_step'],2), round(d['value']), d['phase_ms'])"
