#!/bin/bash
# Round-1 GPU evidence run: parity tests, the bench line, the ncu launch list and --set full captures of the top kernels.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err
tail -c 600 gpurun_out/bench6.err
Q="bench.py --quick --steps 1 --rollout-length 32 --no-cpu-baseline --no-profile"
python $Q > gpurun_out/plain_r1c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/launches_r1c.csv python $Q > gpurun_out/ncu_r1c.log 2>&1
wc -l gpurun_out/launches_r1c.csv
for spec in "gemm_tc_kernel:2200" "gemm_tc_tn_kernel:20" "retention_chunk_fwd_kernel:2" "retention_chunk_bwd_kernel:2" "act_rms_fwd_kernel:1500" "gru_gate_bwd_kernel:40"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s $s -c 2 -f -o gpurun_out/full_r1c_$k python $Q > gpurun_out/ncu_full_$k.log 2>&1
  tail -1 gpurun_out/ncu_full_$k.log
done
ls -la gpurun_out/*.ncu-rep
