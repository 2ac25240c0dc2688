#!/bin/bash
# Round-2 GPU evidence run: ncu launch lists (time + DRAM bytes per launch) of one update minibatch and of a 4-step rollout at the
# bench size, and --set full captures of the kernels named on the command line (default: the fused row-chain kernels).
# Usage (from the repo root, on the GPU box): bash tools/gpu_profile_r2.sh <tag> [kernel:skip ...]
set -u
TAG=${1:-r2}; shift
SPECS=${@:-"chain_gate_kernel:1 chain_tail_kernel:0"}
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv"
python tools/profile_update.py > gpurun_out/pu_plain.log 2>&1 && \
ncu $M --log-file gpurun_out/launches_${TAG}_minibatch.csv python tools/profile_update.py > gpurun_out/pu_ncu.log 2>&1
python tools/profile_update.py --rollout-steps 4 > gpurun_out/pr_plain.log 2>&1 && \
ncu $M --log-file gpurun_out/launches_${TAG}_rollout4.csv python tools/profile_update.py --rollout-steps 4 > gpurun_out/pr_ncu.log 2>&1
for spec in $SPECS; do
  k=${spec%%:*}; s=${spec##*:}
  extra=""; case "$k" in sable_step_kernel|lbf_step_kernel) extra="--rollout-steps 4";; esac
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$k" -s $s -c 1 -f \
    -o gpurun_out/full_${TAG}_$k python tools/profile_update.py $extra > gpurun_out/ncu_full_$k.log 2>&1
  tail -1 gpurun_out/ncu_full_$k.log
done
ls -la gpurun_out/full_${TAG}_*.ncu-rep
