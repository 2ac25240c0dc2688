#!/bin/bash
# Round-2 evidence run (one B200): the GPU test suite, the bench lines of the default workload and of the RWARE shard, ncu launch lists of one
# update minibatch and a 4-step rollout, and --set full captures of the kernels this round built or rebuilt. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/tests_r2_final.txt; cat gpurun_out/tests_r2_final.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_lbf_bench.json 2> gpurun_out/r2_lbf_bench.err; tail -c 200 gpurun_out/r2_lbf_bench.err
python bench.py --env rware --num-envs 1024 --update-batch-size 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_rware_bench.json 2> gpurun_out/r2_rware_bench.err
bash tools/gpu_profile_r2.sh r2 chain_front_kernel:0 chain_gate_kernel:1 chain_tail_kernel:0 gru_scan_fwd_kernel:0 gru_scan_bwd_kernel:0 gemm_tc_kernel:9 gemm_tc_tn_kernel:3 retention_chunk_fwd_kernel:1 sable_step_kernel:1
