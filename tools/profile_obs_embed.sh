for k in obs_embed_fwd_kernel obs_embed_bwd_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$k" -s 0 -c 1 -f -o gpurun_out/full_r1g_$k python tools/profile_update.py > gpurun_out/ncu_full_$k.log 2>&1; tail -1 gpurun_out/ncu_full_$k.log
done
