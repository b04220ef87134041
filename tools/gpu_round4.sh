#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
tag="${1:-it}"
out=gpurun_out; mkdir -p "$out"
python -m pytest tests -m gpu -q -x > "$out/${tag}_pytest.txt" 2>&1
tail -n 6 "$out/${tag}_pytest.txt"
for v in "" "B200NN_U8_TWO_GROUPS=1"; do
  env $v python bench.py --workload cfg2 --steps 20 --warmup 3 > "$out/${tag}_bench_cfg2${v:+_2g}.json" 2> "$out/${tag}_bench_cfg2.err"; tail -c 300 "$out/${tag}_bench_cfg2.err"
  python - "$out/${tag}_bench_cfg2${v:+_2g}.json" "$v" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("cfg2", sys.argv[2], ": qps %.0f ms %.3f frac %.3f" % (l["value"], l["ms_per_step"], l["roofline"]["frac"]), l.get("parity"))
except Exception as e: print("cfg2 unreadable", e)
PY
done
timeout 600 python tests/latency_bench.py > "$out/${tag}_latency.json" 2> "$out/${tag}_latency.err"; grep -E "single_query|batch_ms|k10|k32|k100|64_queries|8_frames" "$out/${tag}_latency.json"; tail -c 300 "$out/${tag}_latency.err"
timeout 120 python tools/quick_ivf_bench.py > "$out/${tag}_ivf_bench.txt" 2>&1; tail -n 4 "$out/${tag}_ivf_bench.txt"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file "$out/${tag}_launches_cfg2.csv" python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > "$out/${tag}_launches_cfg2.log" 2>&1
$NCU --set full --import-source on -k regex:u8_scan_tc -s 15 -c 1 -f -o "$out/${tag}_u8_full" python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline > "$out/${tag}_u8_full.log" 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:pq_encode -c 12 --csv --log-file "$out/${tag}_launches_encode.csv" python tools/ncu_probe.py encode > /dev/null 2>&1
python - "$out/${tag}_launches_cfg2.csv" "$out/${tag}_launches_encode.csv" <<'PY'
import csv,sys
for path in sys.argv[1:]:
    rows=list(csv.reader(open(path))); hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; hdr=rows[hi]
    kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value"); mu=hdr.index("Metric Unit")
    print(path)
    for r in rows[hi+1:][-10:]:
        v=float(r[mv].replace(",","")); u=r[mu]; us=v/1000 if u=="ns" else (v if u=="us" else v*1000)
        print("  %9.1f us  %s" % (us, r[kn].split("(")[0][-50:]))
PY
