#!/usr/bin/env bash
# Second compute-sanitizer pass of a round (see tools/run_sanitizer.sh and profiles/r02_sanitizer_summary.md):
#  * memcheck over the small-size suite again, now including the kernels added after the first pass (tiled fp32 scan, tiled PQ
#    encode, u8 passes, multi-GPU layer on one device, makeSearch drop-in);
#  * racecheck with the kernels that merge into LOCK-protected shared top-k lists excluded (racecheck orders accesses by barriers
#    only and reports those by construction): everything else -- u8 tensor-core scan, merges, tiled kernels, tcgen05 GEMMs, LUT
#    build, training -- must come back with zero hazards.
set -u
cd "$(dirname "$0")/.."
tag="${1:-run}"
out=gpurun_out; mkdir -p "$out"
SEL='not full_size and not cfg5 and not cfg3 and not cfg2 and not cfg4 and not 12p5m and not one_million and not torchrun'
FILES="tests/test_pq_gpu.py tests/test_flat_sq_gpu.py tests/test_frontend_gpu.py tests/test_train_gpu.py tests/test_rotate_gemm_gpu.py tests/test_multi_gpu.py tests/test_makesearch_gpu.py"
rc=0
timeout 1200 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m pytest $FILES -m gpu -q -x -k "$SEL" \
    > "$out/${tag}_sanitizer2_memcheck.log" 2>&1 || rc=$?
echo "memcheck exit code: $rc" | tee -a "$out/${tag}_sanitizer2_memcheck.log"
rc2=0
EXCL="--kernel-name-exclude kns=adc_scan_topk --kernel-name-exclude kns=ivf_search_topk --kernel-name-exclude kns=dense_topk --kernel-name-exclude kns=flat_scan_u8"
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis $EXCL --error-exitcode 9 python -m pytest tests/test_pq_gpu.py tests/test_flat_sq_gpu.py \
    tests/test_rotate_gemm_gpu.py tests/test_frontend_gpu.py -m gpu -q -k "$SEL" > "$out/${tag}_sanitizer2_racecheck.log" 2>&1 || rc2=$?
echo "racecheck (lock-protected list kernels excluded) exit code: $rc2" | tee -a "$out/${tag}_sanitizer2_racecheck.log"
for t in memcheck racecheck; do
    echo "== $t"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" "$out/${tag}_sanitizer2_$t.log" | tail -n 6
done
exit $(( rc | rc2 ))
