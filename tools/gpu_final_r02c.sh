#!/usr/bin/env bash
# Round-2 closing measurement set on ONE GPU.  Order: parity (tests, smoke), ncu captures (exported to CSV on the box; the
# reports themselves are too large to travel), then the bench lines -- the scan's dram traffic of the timed shape is known by then.
set -u
cd "$(dirname "$0")/.."
tag="${1:-r02c}"
out=gpurun_out; mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q > "$out/${tag}_pytest.txt" 2>&1; tail -n 3 "$out/${tag}_pytest.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.txt" 2>&1; tail -n 2 "$out/${tag}_smoke.txt"
NCU="ncu --clock-control none"
full() {  # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o "$out/${tag}_${name}" "$@" > "$out/${tag}_${name}.log" 2>&1
  ncu -i "$out/${tag}_${name}.ncu-rep" --page raw --csv > "$out/${tag}_${name}_ncu_full_raw.csv" 2>/dev/null
}
full scan adc_scan 3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-parity --no-k10
ncu -i "$out/${tag}_scan.ncu-rep" --page source --csv > "$out/${tag}_scan_ncu_source.csv" 2>/dev/null
python tools/ncu_traffic.py "$out/${tag}_scan_ncu_full_raw.csv" 16 1000000 4096 100 "profiles/${tag}_scan_ncu_full_raw.csv (ncu --set full, one launch of bench.py's cfg3 step)"
full scan_m32 adc_scan 2 python tools/quick_scan_bench.py 1000000 4096 32 100
full scan_125k adc_scan 2 python tools/quick_scan_bench.py 125000 4096 16 100
python tools/ncu_traffic.py "$out/${tag}_scan_125k_ncu_full_raw.csv" 16 125000 4096 100 "profiles/${tag}_scan_125k_ncu_full_raw.csv (one GPU's shard of an 8-GPU row-sharded cfg3 step)"
full u8 u8_scan_tc 3 python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline
full ivf ivf_search 2 python tools/ncu_probe.py ivf
full dense_topk dense_topk 3 python tools/ncu_probe.py flat
rm -f "$out"/*.ncu-rep
cp profiles/scan_traffic.json "$out/${tag}_scan_traffic.json"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file "$out/${tag}_launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-parity --no-k10 > "$out/${tag}_launches.log" 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file "$out/${tag}_launches_cfg2.csv" python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > "$out/${tag}_launches_cfg2.log" 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:"flat_|dense_topk|topk_merge|rank_to" -c 60 --csv --log-file "$out/${tag}_launches_flat_makesearch_shape.csv" python tools/ncu_probe.py flat > /dev/null 2>&1
python bench.py --steps 20 --warmup 3 > "$out/${tag}_bench_cfg3_n1.json" 2> "$out/${tag}_bench_cfg3_n1.err"
python bench.py --workload cfg2 --steps 20 --warmup 3 > "$out/${tag}_bench_cfg2.json" 2> "$out/${tag}_bench_cfg2.err"
python bench.py --workload cfg4 --steps 8 --warmup 3 > "$out/${tag}_bench_cfg4_n1.json" 2> "$out/${tag}_bench_cfg4_n1.err"
python bench.py --workload cfg5 --steps 4 --warmup 3 > "$out/${tag}_bench_cfg5_n1.json" 2> "$out/${tag}_bench_cfg5_n1.err"
python bench.py --impl reference --steps 3 --warmup 3 > "$out/${tag}_bench_reference_arm.json" 2> "$out/${tag}_bench_reference_arm.err"
for f in cfg3_n1 cfg2 cfg4_n1 cfg5_n1; do python - "$out/${tag}_bench_$f.json" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "qps %.0f ms %.4f frac %.3f e2e %.0f traffic %s" % (l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"], l["roofline"].get("traffic")), l.get("parity"), l.get("parity_full_scan"), l.get("stage_ms"), l.get("cpu_baseline"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -c 300 "$out/${tag}_bench_reference_arm.json"
timeout 600 python tests/latency_bench.py > "$out/${tag}_latency.json" 2> "$out/${tag}_latency.err"; grep -E "single_query|batch_ms|k10|k32|k100|64_queries|8_frames|ms\"" "$out/${tag}_latency.json" | head -30
timeout 300 python tools/quick_ivf_bench.py > "$out/${tag}_ivf_bench.txt" 2>&1; tail -n 4 "$out/${tag}_ivf_bench.txt"
du -sh "$out"
