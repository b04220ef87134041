#!/usr/bin/env bash
# cfg4-shaped scan (M = 32, G = 8 lane groups): both ring variants timed, candidate counters, one ncu capture each
set -u
cd "$(dirname "$0")/.."
tag="${1:-g8}"
out=gpurun_out; mkdir -p "$out"
for v in 0 1; do
  B200NN_SCAN_VAR=$v QUICK_STATS=1 timeout 300 python tools/quick_scan_bench.py 1000000 4096 32 100 > "$out/${tag}_var${v}.txt" 2>&1
  grep -E "iter [1-4]|sha1|scan stats" "$out/${tag}_var${v}.txt"
done
NCU="ncu --clock-control none"
for v in 0 1; do
  B200NN_SCAN_VAR=$v timeout 600 $NCU --set full --import-source on -k regex:adc_scan -s 2 -c 1 -f -o "$out/${tag}_var${v}_full" python tools/quick_scan_bench.py 1000000 4096 32 100 > "$out/${tag}_var${v}_ncu.log" 2>&1
done
ls -la $out | grep "${tag}_"
