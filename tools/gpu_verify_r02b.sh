#!/usr/bin/env bash
# round-2 second measurement set on ONE GPU: tests, the four single-GPU contract lines, launch list + full ncu captures
set -u
cd "$(dirname "$0")/.."
tag="${1:-r02b}"
out=gpurun_out; mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q > "$out/${tag}_pytest.txt" 2>&1; tail -n 3 "$out/${tag}_pytest.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.txt" 2>&1; tail -n 2 "$out/${tag}_smoke.txt"
python bench.py --steps 20 --warmup 3 > "$out/${tag}_bench_cfg3_n1.json" 2> "$out/${tag}_bench_cfg3_n1.err"
python bench.py --workload cfg2 --steps 20 --warmup 3 > "$out/${tag}_bench_cfg2.json" 2> "$out/${tag}_bench_cfg2.err"
python bench.py --workload cfg4 --steps 8 --warmup 3 > "$out/${tag}_bench_cfg4_n1.json" 2> "$out/${tag}_bench_cfg4_n1.err"
python bench.py --workload cfg5 --steps 5 --warmup 3 > "$out/${tag}_bench_cfg5_n1.json" 2> "$out/${tag}_bench_cfg5_n1.err"
for f in cfg3_n1 cfg2 cfg4_n1 cfg5_n1; do python - "$out/${tag}_bench_$f.json" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "qps %.0f ms %.4f frac %.3f e2e %.0f" % (l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"]), l.get("parity"), l.get("parity_full_scan"), l.get("stage_ms"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file "$out/${tag}_launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-parity --no-k10 > "$out/${tag}_launches.log" 2>&1
$NCU --set full --import-source on -k regex:adc_scan -s 3 -c 1 -f -o "$out/${tag}_scan_full" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-parity --no-k10 > "$out/${tag}_scan_full.log" 2>&1
$NCU --set full --import-source on -k regex:u8_scan_tc -s 3 -c 1 -f -o "$out/${tag}_u8_full" python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline > "$out/${tag}_u8_full.log" 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file "$out/${tag}_launches_cfg2.csv" python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > "$out/${tag}_launches_cfg2.log" 2>&1
$NCU --set full --import-source on -k regex:adc_scan -s 2 -c 1 -f -o "$out/${tag}_scan_m32_full" python tools/quick_scan_bench.py 1000000 4096 32 100 > "$out/${tag}_scan_m32_full.log" 2>&1
$NCU --set full --import-source on -k regex:adc_scan -s 2 -c 1 -f -o "$out/${tag}_scan_125k_full" python tools/quick_scan_bench.py 125000 4096 16 100 > "$out/${tag}_scan_125k_full.log" 2>&1
ls -la "$out" | grep "${tag}_" | awk '{print $5, $9}'
