#!/usr/bin/env bash
# development aid: parity suite + secondary benches of one iteration on ONE GPU
set -u
cd "$(dirname "$0")/.."
tag="${1:-it}"
out=gpurun_out; mkdir -p "$out"
python -m pytest tests -m gpu -q -x > "$out/${tag}_pytest.txt" 2>&1
tail -n 6 "$out/${tag}_pytest.txt"
tools/bin/i8_peak > "$out/${tag}_i8_peak.json" 2>&1; cat "$out/${tag}_i8_peak.json"
python bench.py --workload cfg2 --steps 20 --warmup 3 > "$out/${tag}_bench_cfg2.json" 2> "$out/${tag}_bench_cfg2.err"; tail -c 300 "$out/${tag}_bench_cfg2.err"
python - "$out/${tag}_bench_cfg2.json" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("cfg2: qps %.0f ms %.3f frac %.3f" % (l["value"], l["ms_per_step"], l["roofline"]["frac"]), l.get("parity"))
except Exception as e: print("cfg2 unreadable", e)
PY
timeout 600 python tests/latency_bench.py > "$out/${tag}_latency.json" 2> "$out/${tag}_latency.err"; cat "$out/${tag}_latency.json"; tail -c 300 "$out/${tag}_latency.err"
timeout 120 python tools/quick_ivf_bench.py > "$out/${tag}_ivf_bench.txt" 2>&1; tail -n 6 "$out/${tag}_ivf_bench.txt"
