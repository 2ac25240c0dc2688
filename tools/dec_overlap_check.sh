timeout 400 python -m pytest tests/test_gpu_networks.py tests/test_gpu_learner.py tests/test_gpu_lbf.py tests/test_golden_traces.py -q -m gpu -x 2>&1 | tail -2
for m in 0 1; do
MAGPO_DEC_OVERLAP=$m timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_dec${m}_lbf.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_dec${m}_lbf.json')); print('dec_overlap=${m} lbf', round(d['ms_per_step'],2), d['phase_ms'])"
done
MAGPO_DEC_OVERLAP=1 timeout 200 python bench.py --env rware --num-envs 1024 --update-batch-size 1 --steps 4 --warmup 3 --no-cpu-baseline --no-profile 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rware dec1', d['ms_per_step'], d['phase_ms'])"
