#!/usr/bin/env bash
# the racecheck half of tools/run_sanitizer2.sh alone
set -u
cd "$(dirname "$0")/.."
tag="${1:-run}"; out=gpurun_out; mkdir -p "$out"
SEL='not full_size and not cfg5 and not cfg3 and not cfg2 and not cfg4 and not 12p5m and not one_million and not torchrun'
EXCL="--kernel-name-exclude kns=adc_scan_topk --kernel-name-exclude kns=ivf_search_topk --kernel-name-exclude kns=dense_topk --kernel-name-exclude kns=flat_scan_u8"
rc2=0
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis $EXCL --error-exitcode 9 python -m pytest tests/test_pq_gpu.py tests/test_flat_sq_gpu.py \
    tests/test_rotate_gemm_gpu.py tests/test_frontend_gpu.py -m gpu -q -k "$SEL" > "$out/${tag}_sanitizer2_racecheck.log" 2>&1 || rc2=$?
echo "racecheck (lock-protected list kernels excluded) exit code: $rc2" | tee -a "$out/${tag}_sanitizer2_racecheck.log"
grep -E "RACECHECK SUMMARY|passed|failed|exit code" "$out/${tag}_sanitizer2_racecheck.log" | tail -n 5
grep -E "Race reported" "$out/${tag}_sanitizer2_racecheck.log" | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -20
