#!/usr/bin/env bash
# iteration check: full GPU test suite, scan variants at the three shapes, cfg2 with and without the shared bound
set -u
cd "$(dirname "$0")/.."
tag="${1:-it3}"
out=gpurun_out; mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -x -q > "$out/${tag}_pytest.txt" 2>&1; tail -n 4 "$out/${tag}_pytest.txt"
run() {  # rows batch M var
  B200NN_SCAN_VAR=$4 QUICK_STATS=1 timeout 300 python tools/quick_scan_bench.py $1 $2 $3 100 > "$out/${tag}_n$1_m$3_var$4.txt" 2>&1
  echo "== rows $1 batch $2 M $3 var $4"; grep -E "iter [3-4]|sha1|scan stats" "$out/${tag}_n$1_m$3_var$4.txt" | sed 's/.*lut_ms/lut_ms/' | cut -c1-420
}
run 1000000 4096 16 2
for v in 0 2; do run 125000 4096 16 $v; done
run 500000 1024 16 0; run 500000 1024 16 2
run 1000000 4096 32 3
for v in "" "B200NN_U8_NO_SHARED_BOUND=1"; do
  env $v timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > "$out/${tag}_bench_cfg2${v:+_3pass}.json" 2> "$out/${tag}_bench_cfg2${v:+_3pass}.err"
  python - "$out/${tag}_bench_cfg2${v:+_3pass}.json" "$v" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("cfg2", sys.argv[2], ": qps %.0f ms %.4f frac %.3f e2e %.0f" % (l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"]), l.get("parity"))
except Exception as e: print("cfg2 unreadable", e)
PY
done
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file "$out/${tag}_launches_cfg2.csv" python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > "$out/${tag}_launches_cfg2.log" 2>&1
tail -n 12 "$out/${tag}_launches_cfg2.csv" | cut -d, -f5,12- | cut -c1-200
