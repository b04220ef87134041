#!/usr/bin/env bash
# development aid: parity suite + the benches of one iteration on ONE GPU; everything lands in gpurun_out/<tag>_*
set -u
cd "$(dirname "$0")/.."
tag="${1:-it}"
out=gpurun_out
mkdir -p "$out"
python -m pytest tests -m gpu -q -x > "$out/${tag}_pytest.txt" 2>&1
tail -n 5 "$out/${tag}_pytest.txt"
python bench.py --steps 10 --warmup 3 > "$out/${tag}_bench_cfg3.json" 2> "$out/${tag}_bench_cfg3.err"; tail -c 600 "$out/${tag}_bench_cfg3.err"
python bench.py --workload cfg5_small --steps 5 --warmup 3 > "$out/${tag}_bench_cfg5small.json" 2> "$out/${tag}_bench_cfg5small.err"; tail -c 600 "$out/${tag}_bench_cfg5small.err"
python bench.py --workload cfg4_shard --steps 5 --warmup 3 > "$out/${tag}_bench_cfg4shard.json" 2> "$out/${tag}_bench_cfg4shard.err"; tail -c 600 "$out/${tag}_bench_cfg4shard.err"
for f in cfg3 cfg5small cfg4shard; do python - "$out/${tag}_bench_$f.json" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "qps %.0f ms %.3f frac %.3f e2e %.0f" % (l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"]), l.get("parity"), l.get("parity_full_scan"), l.get("sanity"), l.get("k10"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
