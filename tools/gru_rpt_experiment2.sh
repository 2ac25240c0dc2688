for r in 128 112 96 80; do
  MAGPO_GRU_RPT=$r timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_rpt2_${r}_lbf.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_rpt2_${r}_lbf.json")); print("rpt=${r}", round(d["ms_per_step"],2), d["phase_ms"])
PY
done
