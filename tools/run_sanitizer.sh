#!/usr/bin/env bash
# CI-style memory / race check of the CUDA library (the reference has no sanitizer coverage at all).
# Runs the small-size parity tests -- every kernel family is launched by them -- under compute-sanitizer on ONE GPU:
#     gpurun --timeout 1500 -- 'bash tools/run_sanitizer.sh > gpurun_out/sanitizer.log 2>&1'
# memcheck on everything small; racecheck + synccheck on the kernels that share memory between warps (fused scan, top-k merge,
# tcgen05 kernels).  The full-size tests are excluded (the tools slow kernels down by 10-100x).
# NOT run in round 1 (the round's GPU budget went to parity runs, benchmarks and ncu); results belong in profiles/.
set -u
cd "$(dirname "$0")/.."
SEL='not full_size and not cfg5 and not cfg3 and not cfg2 and not cfg4'
FILES="tests/test_pq_gpu.py tests/test_flat_sq_gpu.py tests/test_frontend_gpu.py tests/test_train_gpu.py tests/test_rotate_gemm_gpu.py"
rc=0
compute-sanitizer --tool memcheck --error-exitcode 9 --leak-check full python -m pytest $FILES -m gpu -q -x -k "$SEL" || rc=$?
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_pq_gpu.py tests/test_flat_sq_gpu.py -m gpu -q -x \
    -k "golden or tail_pieces or clamp_ties or tensor_core" || rc=$?
compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_pq_gpu.py tests/test_rotate_gemm_gpu.py -m gpu -q -x \
    -k "golden or bit" || rc=$?
echo "sanitizer exit code: $rc"
exit $rc
