#!/bin/bash
# compute-sanitizer pass over smoke() (one small rollout + update on the LBF workload, every kernel of the path incl. the tcgen05 / TMA
# pipelines). Usage (GPU box): bash tools/gpu_sanitize.sh memcheck|racecheck|synccheck|initcheck   -> gpurun_out/sanitizer_<tool>.txt
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$TOOL.txt 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_$TOOL.txt
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|exit code|Error|hazard" gpurun_out/sanitizer_$TOOL.txt | sort | uniq -c | sort -rn | head -20
