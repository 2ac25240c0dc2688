for c in 2048 1024; do
timeout 200 python bench.py --chunk-envs $c --steps 4 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_chunk${c}_lbf.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_chunk${c}_lbf.json')); print('chunk=${c} lbf', round(d['ms_per_step'],2), d['phase_ms'])"
done
