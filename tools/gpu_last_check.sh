#!/usr/bin/env bash
# last check of the round on ONE GPU: the whole GPU suite on the final code, cfg2 after the per-chunk bound read
set -u
cd "$(dirname "$0")/.."
tag="${1:-last}"
out=gpurun_out; mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q > "$out/${tag}_pytest.txt" 2>&1; tail -n 3 "$out/${tag}_pytest.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.txt" 2>&1; tail -n 1 "$out/${tag}_smoke.txt"
python bench.py --workload cfg2 --steps 20 --warmup 3 > "$out/${tag}_bench_cfg2.json" 2> "$out/${tag}_bench_cfg2.err"
python - "$out/${tag}_bench_cfg2.json" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("cfg2: qps %.0f ms %.4f frac %.3f e2e %.0f" % (l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"]), l.get("parity"))
PY
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-parity > "$out/${tag}_bench_cfg3_quick.json" 2> "$out/${tag}_bench_cfg3_quick.err"
python - "$out/${tag}_bench_cfg3_quick.json" <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("cfg3: qps %.0f ms %.4f frac %.3f e2e %.0f" % (l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"]))
PY
