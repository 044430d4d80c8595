#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run on a B200 box: gpurun -- 'bash tools/sanitize.sh').
# memcheck: out-of-bounds / misaligned global + shared accesses over the whole parity + trainer suites;
# racecheck / synccheck: shared-memory hazards and barrier misuse in the warp-specialised mbarrier / TMEM kernels
# (v4 forward, v3 forward, fused backward) -- selected with -k because racecheck slows those kernels ~100x.
# Each pass is bounded; logs land in gpurun_out/sanitize_<tool>.log, one-line verdicts in gpurun_out/sanitize_summary.txt.
# KGCN_PDL=0: the sanitizer serialises launches anyway and does not model programmatic dependent launch.
set -u
mkdir -p gpurun_out
ALL="tests/test_gpu_parity.py tests/test_gpu_trainer.py tests/test_next_rows.py"
FUSED="v4_pipeline or backward_fused or full_size or step_matches_oracle or chained_launches or multi_step_graph or graph_replay"
: > gpurun_out/sanitize_summary.txt
run() {   # tool, timeout, pytest args...
    local tool=$1 tmo=$2; shift 2
    KGCN_PDL=0 timeout "$tmo" compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
        python -m pytest "$@" -m gpu -q -p no:cacheprovider > "gpurun_out/sanitize_${tool}.log" 2>&1
    local rc=$?
    echo "$tool: exit $rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}.log | tail -1); pytest: $(grep -E 'passed|failed' gpurun_out/sanitize_${tool}.log | tail -1)" | tee -a gpurun_out/sanitize_summary.txt
}
run memcheck "${SANITIZE_TIMEOUT:-300}" $ALL
run racecheck "${SANITIZE_TIMEOUT:-300}" tests/test_gpu_parity.py tests/test_gpu_trainer.py -k "$FUSED"
run synccheck "${SANITIZE_TIMEOUT:-240}" tests/test_gpu_parity.py tests/test_gpu_trainer.py -k "$FUSED"
