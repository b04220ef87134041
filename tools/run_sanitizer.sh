#!/usr/bin/env bash
# CI-style memory / race check of the CUDA library (the reference has no sanitizer coverage at all).
# Runs the small-size parity tests -- every kernel family is launched by them -- under compute-sanitizer on ONE GPU:
#     gpurun --timeout 1500 -- 'bash tools/run_sanitizer.sh r02'
# memcheck on everything small; racecheck + synccheck on the kernels that share memory between warps (fused scan, top-k merge,
# tcgen05 kernels).  The full-size tests are excluded (the tools slow kernels down by 10-100x).
# Logs: gpurun_out/<tag>_sanitizer_{memcheck,racecheck,synccheck}.log (copied to profiles/ once read).
set -u
cd "$(dirname "$0")/.."
tag="${1:-run}"
out=gpurun_out
mkdir -p "$out"
SEL='not full_size and not cfg5 and not cfg3 and not cfg2 and not cfg4 and not 12p5m and not one_million'
FILES="tests/test_pq_gpu.py tests/test_flat_sq_gpu.py tests/test_frontend_gpu.py tests/test_train_gpu.py tests/test_rotate_gemm_gpu.py"
rc=0
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $FILES -m gpu -q -x -k "$SEL" \
    > "$out/${tag}_sanitizer_memcheck.log" 2>&1 || rc=$?
echo "memcheck exit code: $rc" | tee -a "$out/${tag}_sanitizer_memcheck.log"
rc2=0
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_pq_gpu.py tests/test_flat_sq_gpu.py -m gpu -q \
    -k "golden or tail_pieces or clamp_ties or tensor_core" > "$out/${tag}_sanitizer_racecheck.log" 2>&1 || rc2=$?
echo "racecheck exit code: $rc2" | tee -a "$out/${tag}_sanitizer_racecheck.log"
rc3=0
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_pq_gpu.py tests/test_rotate_gemm_gpu.py -m gpu -q \
    -k "golden or bit" > "$out/${tag}_sanitizer_synccheck.log" 2>&1 || rc3=$?
echo "synccheck exit code: $rc3" | tee -a "$out/${tag}_sanitizer_synccheck.log"
for t in memcheck racecheck synccheck; do
    echo "== $t"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" "$out/${tag}_sanitizer_$t.log" | tail -n 6
done
exit $(( rc | rc2 | rc3 ))
