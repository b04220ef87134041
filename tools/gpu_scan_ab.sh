#!/usr/bin/env bash
# A/B of the scan kernel variants (B200NN_SCAN_VAR: bit 0 two-slot ring for G = 8, bit 1 grouped threshold check)
set -u
cd "$(dirname "$0")/.."
tag="${1:-ab}"
out=gpurun_out; mkdir -p "$out"
timeout 900 python -m pytest tests/test_pq_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > "$out/${tag}_pytest.txt" 2>&1; tail -n 3 "$out/${tag}_pytest.txt"
run() {  # rows batch M var
  B200NN_SCAN_VAR=$4 QUICK_STATS=1 timeout 300 python tools/quick_scan_bench.py $1 $2 $3 100 > "$out/${tag}_n$1_m$3_var$4.txt" 2>&1
  echo "== rows $1 batch $2 M $3 var $4"; grep -E "iter [2-4]|sha1|scan stats" "$out/${tag}_n$1_m$3_var$4.txt" | sed 's/.*scan_ms/scan_ms/' | cut -c1-420
}
for v in 0 2; do run 1000000 4096 16 $v; done
for v in 0 2; do run 125000 4096 16 $v; done
for v in 0 1 2 3; do run 1000000 4096 32 $v; done
