#!/bin/bash
# Round-1 GPU evidence run: parity tests, the bench line, ncu launch lists (time + DRAM bytes per launch) of one update
# minibatch and of a 4-step rollout at the bench size, and --set full captures of the top kernels.
# Usage (from the repo root, on the GPU box): bash tools/gpu_profile_r1.sh <tag>
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 300 gpurun_out/bench_$TAG.err
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv"
python tools/profile_update.py > gpurun_out/pu_plain.log 2>&1 && \
ncu $M --log-file gpurun_out/launches_${TAG}_minibatch.csv python tools/profile_update.py > gpurun_out/pu_ncu.log 2>&1
python tools/profile_update.py --rollout-steps 4 > gpurun_out/pr_plain.log 2>&1 && \
ncu $M --log-file gpurun_out/launches_${TAG}_rollout4.csv python tools/profile_update.py --rollout-steps 4 > gpurun_out/pr_ncu.log 2>&1
for spec in "gemm_tc_kernel:12" "gemm_tc_tn_kernel:3" "retention_chunk_bwd_kernel:0" "retention_chunk_fwd_kernel:1" "gru_scan_bwd_kernel:0" "gru_scan_fwd_kernel:0" "act_rms_bwd_kernel:1" "thin_n_bwd_kernel:0"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$k" -s $s -c 1 -f \
    -o gpurun_out/full_${TAG}_$k python tools/profile_update.py > gpurun_out/ncu_full_$k.log 2>&1
  tail -1 gpurun_out/ncu_full_$k.log
done
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sable_step_kernel" -s 1 -c 1 -f \
  -o gpurun_out/full_${TAG}_sable_step_kernel python tools/profile_update.py --rollout-steps 4 > gpurun_out/ncu_full_retention_fwd.log 2>&1
tail -1 gpurun_out/ncu_full_retention_fwd.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"lbf_step_kernel" -s 1 -c 1 -f \
  -o gpurun_out/full_${TAG}_lbf_step_kernel python tools/profile_update.py --rollout-steps 4 > gpurun_out/ncu_full_lbf_step.log 2>&1
tail -1 gpurun_out/ncu_full_lbf_step.log
python tools/profile_update.py --gae > gpurun_out/pg_plain.log 2>&1 && \
ncu $M --log-file gpurun_out/launches_${TAG}_gae.csv python tools/profile_update.py --gae > gpurun_out/pg_ncu.log 2>&1
ls -la gpurun_out/full_${TAG}_*.ncu-rep
