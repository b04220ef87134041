#!/usr/bin/env bash
# Third compute-sanitizer pass (second half of round 2): the kernels changed or added since the second pass --
# grouped-test ADC scan + two-slot ring, one-pass u8 scan with the bound warp, IVF search and dense selection on the CTA
# buffer (cta_sort_trim: barrier-ordered only, so racecheck now INCLUDES them), LUT build.
#   memcheck : the whole small-size suite
#   synccheck: the selection kernels and the scan (divergent barriers are exactly what a non-uniform sort decision would cause)
#   racecheck: everything except the two kernels that merge under locks (adc_scan_topk, flat_scan_u8)
set -u
cd "$(dirname "$0")/.."
tag="${1:-run}"
out=gpurun_out; mkdir -p "$out"
SEL='not full_size and not cfg5 and not cfg3 and not cfg2 and not cfg4 and not 12p5m and not one_million and not torchrun'
FILES="tests/test_pq_gpu.py tests/test_flat_sq_gpu.py tests/test_frontend_gpu.py tests/test_train_gpu.py tests/test_rotate_gemm_gpu.py tests/test_multi_gpu.py tests/test_makesearch_gpu.py"
rc=0
timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m pytest $FILES -m gpu -q -x -k "$SEL" \
    > "$out/${tag}_sanitizer3_memcheck.log" 2>&1 || rc=$?
echo "memcheck exit code: $rc" | tee -a "$out/${tag}_sanitizer3_memcheck.log"
rc3=0
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_pq_gpu.py tests/test_flat_sq_gpu.py tests/test_makesearch_gpu.py -m gpu -q \
    -k "ivf or long_lists or variants or golden or query_groups or flat_vs or flat_larger or makesearch or tail_pieces" > "$out/${tag}_sanitizer3_synccheck.log" 2>&1 || rc3=$?
echo "synccheck exit code: $rc3" | tee -a "$out/${tag}_sanitizer3_synccheck.log"
rc2=0
EXCL="--kernel-name-exclude kns=adc_scan_topk --kernel-name-exclude kns=flat_scan_u8"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis $EXCL --error-exitcode 9 python -m pytest tests/test_pq_gpu.py tests/test_flat_sq_gpu.py \
    tests/test_makesearch_gpu.py -m gpu -q -k "($SEL) and (ivf or long_lists or golden or query_groups or flat_vs or flat_larger or makesearch or tensor_core or K8192)" \
    > "$out/${tag}_sanitizer3_racecheck.log" 2>&1 || rc2=$?
echo "racecheck (lock-merging kernels excluded) exit code: $rc2" | tee -a "$out/${tag}_sanitizer3_racecheck.log"
for t in memcheck synccheck racecheck; do
    echo "== $t"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" "$out/${tag}_sanitizer3_$t.log" | tail -n 6
done
grep -E "Race reported|Barrier error|Invalid" "$out/${tag}_sanitizer3_racecheck.log" "$out/${tag}_sanitizer3_synccheck.log" "$out/${tag}_sanitizer3_memcheck.log" | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -20
exit 0
