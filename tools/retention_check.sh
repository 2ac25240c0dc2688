timeout 300 python -m pytest tests/test_gpu_primitives.py -q -m gpu -k retention 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_learner.py tests/test_gpu_lbf.py tests/test_gpu_networks.py -q -m gpu -x 2>&1 | tail -2
timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ldsm_lbf.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_ldsm_lbf.json')); b=d['breakdown_ms_per_step']; print('lbf', round(d['ms_per_step'],2), d['phase_ms'], b['retention_fwd']['ms'], b['retention_bwd']['ms'])"
