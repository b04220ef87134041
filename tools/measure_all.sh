#!/usr/bin/env bash
# One-call measurement set for a round (development aid): everything lands in gpurun_out/<tag>_*.
#     gpurun --timeout 900 -- 'bash tools/measure_all.sh r02'
# 1 GPU.  Order: parity first (a fast kernel whose results differ is not done), then the contract bench, the secondary
# workloads, the launch list and ONE full ncu capture of the dominant kernel.  Numbers printed under ncu are never bench values.
set -u
cd "$(dirname "$0")/.."
tag="${1:-run}"
out=gpurun_out
mkdir -p "$out"
python -m pytest tests -m gpu -q > "$out/${tag}_pytest.txt" 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.txt" 2>&1
python bench.py --steps 20 --warmup 3 > "$out/${tag}_bench_n1.json" 2> "$out/${tag}_bench_n1.err"
python bench.py --workload cfg2 --steps 20 --warmup 3 > "$out/${tag}_bench_cfg2.json" 2> "$out/${tag}_bench_cfg2.err"
timeout 120 python tools/quick_front_bench.py > "$out/${tag}_front_bench.txt" 2>&1
timeout 200 python tools/quick_train_bench.py > "$out/${tag}_train_bench.txt" 2>&1
timeout 200 python tools/quick_ivf_bench.py > "$out/${tag}_ivf_bench.txt" 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/${tag}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$out/${tag}_launches.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:adc_scan -s 3 -c 1 -o "$out/${tag}_scan_full" \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > "$out/${tag}_scan_full.log" 2>&1
tail -n 3 "$out/${tag}_pytest.txt"
cat "$out/${tag}_smoke.txt"
