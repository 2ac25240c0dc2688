#!/bin/bash
# c4 bench line (RWARE small-4ag, long rollouts), 5-seed LBF training curves, curve parity against the oracle on configs[0].
mkdir -p gpurun_out
timeout 300 python bench.py --env rware-small --num-envs 1024 --update-batch-size 1 --rollout-length 512 --chunk-envs 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_rware_small.json 2> gpurun_out/bench_c4_rware_small.err; head -c 200 gpurun_out/bench_c4_rware_small.json; tail -2 gpurun_out/bench_c4_rware_small.err; echo
for seed in 0 1 2 3 4; do
  timeout 300 python -m magpo_b200.rec_magpo env=lbf arch.num_envs=1024 system.num_updates=160 arch.num_evaluation=8 system.total_timesteps=~ system.seed=$seed arch.absolute_metric=False > gpurun_out/train_lbf_seed$seed.log 2>&1; tail -1 gpurun_out/train_lbf_seed$seed.log
done
timeout 900 python tools/curve_parity.py --seeds 5 --updates 6 --out gpurun_out/curve_parity_c1.json > gpurun_out/curve_parity_c1.log 2>&1; tail -3 gpurun_out/curve_parity_c1.log
