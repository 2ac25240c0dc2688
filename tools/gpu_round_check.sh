#!/bin/bash
# Parity suite + bench lines + a short LBF training run; outputs under gpurun_out/. Usage: bash tools/gpu_round_check.sh <tag>
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/tests_$TAG.txt; tail -3 gpurun_out/tests_$TAG.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_lbf.json 2> gpurun_out/bench_${TAG}_lbf.err; head -c 200 gpurun_out/bench_${TAG}_lbf.json; echo
timeout 300 python bench.py --env rware --num-envs 1024 --update-batch-size 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_rware.json 2> gpurun_out/bench_${TAG}_rware.err; head -c 200 gpurun_out/bench_${TAG}_rware.json; echo
if [ "${2:-}" = "train" ]; then
  timeout 400 python -m magpo_b200.rec_magpo env=lbf arch.num_envs=1024 system.num_updates=300 arch.num_evaluation=10 system.total_timesteps=~ > gpurun_out/train_lbf_seed42.log 2>&1; tail -12 gpurun_out/train_lbf_seed42.log
fi
