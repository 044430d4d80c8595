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
# The wide-layer kernel (graphconv_fused_v5.cu) hands its row sums from the aggregation warps to the epilogue warps through
# shared memory, ordered by the mbarrier chain zfull -> tcgen05.commit -> tfull, which racecheck does not follow (it reports the
# pair as a hazard and then stops recording at its hazard cap): it gets its own pass (suffix _v5), the other kernels are checked
# with KGCN_FUSED_V5=0 so that one kernel's reports cannot hide another's.
run2() {   # suffix, env assignment, tool, timeout, pytest args...
    local sfx=$1 envs=$2 tool=$3 tmo=$4; shift 4
    env KGCN_PDL=0 $envs timeout "$tmo" compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
        python -m pytest "$@" -m gpu -q -p no:cacheprovider > "gpurun_out/sanitize_${tool}${sfx}.log" 2>&1
    local rc=$?
    echo "$tool$sfx ($envs): exit $rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}${sfx}.log | tail -1); pytest: $(grep -E 'passed|failed' gpurun_out/sanitize_${tool}${sfx}.log | tail -1)" | tee -a gpurun_out/sanitize_summary.txt
}
V5="128-128 or full_size"
run2 "" "KGCN_FUSED_V5=1" memcheck "${SANITIZE_TIMEOUT:-300}" $ALL
run2 "" "KGCN_FUSED_V5=0" racecheck "${SANITIZE_TIMEOUT:-300}" tests/test_gpu_parity.py tests/test_gpu_trainer.py -k "$FUSED"
run2 "" "KGCN_FUSED_V5=0" synccheck "${SANITIZE_TIMEOUT:-240}" tests/test_gpu_parity.py tests/test_gpu_trainer.py -k "$FUSED"
run2 "_v5" "KGCN_FUSED_V5=1" synccheck "${SANITIZE_TIMEOUT:-240}" tests/test_gpu_parity.py -k "$V5"
# (the v5 step chain -- fused head, hard job boundaries on a named barrier -- goes through memcheck and racecheck; synccheck would
# stop at the named role barrier, see profiles/r02b_sanitizer.txt)
run2 "_v5" "KGCN_FUSED_V5=1 NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=100000" racecheck "${SANITIZE_TIMEOUT:-300}" tests/test_gpu_parity.py tests/test_gpu_trainer.py -k "$V5 or (step_matches_oracle and 64-1-128)"
