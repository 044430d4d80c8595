#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run on a B200 box: gpurun -- 'bash tools/sanitize.sh').
# memcheck: out-of-bounds / misaligned global + shared accesses; racecheck: shared-memory hazards between the roles of the
# fused kernels; synccheck: barrier misuse.  Each pass is bounded; the logs land in gpurun_out/sanitize_<tool>.log.
# KGCN_PDL=0: the sanitizer serialises launches anyway and does not model programmatic dependent launch.
set -u
mkdir -p gpurun_out
SEL="${1:-tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_next_rows.py}"
for tool in memcheck racecheck synccheck; do
    KGCN_PDL=0 timeout "${SANITIZE_TIMEOUT:-420}" compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
        python -m pytest $SEL -m gpu -x -q -p no:cacheprovider > "gpurun_out/sanitize_${tool}.log" 2>&1
    echo "$tool: exit $? ($(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${tool}.log) summaries)"
    tail -3 "gpurun_out/sanitize_${tool}.log"
done
