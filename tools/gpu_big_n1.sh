#!/usr/bin/env bash
# the two big contract configs on ONE GPU (N = 1 lines of the cfg4 / cfg5 scaling tables)
set -u
cd "$(dirname "$0")/.."
tag="${1:-big}"
out=gpurun_out; mkdir -p "$out"
python bench.py --workload cfg5 --steps 5 --warmup 3 > "$out/${tag}_bench_cfg5_n1.json" 2> "$out/${tag}_bench_cfg5_n1.err"; tail -c 400 "$out/${tag}_bench_cfg5_n1.err"
python bench.py --workload cfg4 --steps 8 --warmup 3 > "$out/${tag}_bench_cfg4_n1.json" 2> "$out/${tag}_bench_cfg4_n1.err"; tail -c 400 "$out/${tag}_bench_cfg4_n1.err"
python bench.py --impl reference --workload cfg5 --steps 2 --warmup 3 > "$out/${tag}_bench_cfg5_reference.json" 2> "$out/${tag}_bench_cfg5_reference.err"
for f in cfg5_n1 cfg4_n1; do python - "$out/${tag}_bench_$f.json" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "qps %.0f ms %.3f frac %.3f e2e %.0f build %.1fs" % (l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"], l["config"]["index_build_s"]), l.get("parity"), l.get("parity_full_scan"), l.get("sanity"), l.get("k10"), l.get("cpu_baseline"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
cat "$out/${tag}_bench_cfg5_reference.json" | cut -c1-400
