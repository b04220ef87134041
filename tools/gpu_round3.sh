#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
tag="${1:-it}"
out=gpurun_out; mkdir -p "$out"
python -m pytest tests -m gpu -q -x > "$out/${tag}_pytest.txt" 2>&1
tail -n 6 "$out/${tag}_pytest.txt"
timeout 600 python tests/latency_bench.py > "$out/${tag}_latency.json" 2> "$out/${tag}_latency.err"; grep -E "single_query|batch_ms|k10|k32|k100|64_queries|8_frames" "$out/${tag}_latency.json"; tail -c 300 "$out/${tag}_latency.err"
timeout 120 python tools/quick_ivf_bench.py > "$out/${tag}_ivf_bench.txt" 2>&1; tail -n 4 "$out/${tag}_ivf_bench.txt"
bash tools/ncu_round.sh "$tag"
