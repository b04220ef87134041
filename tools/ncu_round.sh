#!/usr/bin/env bash
# ncu evidence of a round on ONE GPU (numbers printed under ncu are never bench values): launch list of the contract bench,
# full captures of the dominant kernels; reports land in gpurun_out/<tag>_*.ncu-rep and are exported to CSV here afterwards.
set -u
cd "$(dirname "$0")/.."
tag="${1:-ncu}"
out=gpurun_out; mkdir -p "$out"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file "$out/${tag}_launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-parity --no-k10 > "$out/${tag}_launches.log" 2>&1
$NCU --set full --import-source on -k regex:adc_scan -s 3 -c 1 -f -o "$out/${tag}_scan_full" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-full-parity --no-k10 > "$out/${tag}_scan_full.log" 2>&1
# cfg2: 3 scan launches per search (16 k rows, 112 k rows, the rest): the third one of the 4th search is the bulk pass
$NCU --set full --import-source on -k regex:u8_scan_tc -s 11 -c 1 -f -o "$out/${tag}_u8_full" python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline > "$out/${tag}_u8_full.log" 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file "$out/${tag}_launches_cfg2.csv" python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > "$out/${tag}_launches_cfg2.log" 2>&1
$NCU --set full --import-source on -k regex:flat_tile_f32 -s 2 -c 1 -f -o "$out/${tag}_flat_tile_full" python tools/ncu_probe.py flat > "$out/${tag}_flat_full.log" 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:"flat_|dense_topk|topk_merge|rank_to" -c 60 --csv --log-file "$out/${tag}_launches_flat.csv" python tools/ncu_probe.py flat > /dev/null 2>&1
$NCU --set full --import-source on -k regex:pq_encode_tile -s 2 -c 1 -f -o "$out/${tag}_encode_full" python tools/ncu_probe.py encode > "$out/${tag}_encode_full.log" 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:pq_encode -c 12 --csv --log-file "$out/${tag}_launches_encode.csv" python tools/ncu_probe.py encode > /dev/null 2>&1
$NCU --set full --import-source on -k regex:ivf_search_topk -s 2 -c 1 -f -o "$out/${tag}_ivf_full" python tools/ncu_probe.py ivf > "$out/${tag}_ivf_full.log" 2>&1
ls -la "$out" | grep "${tag}_" | head -30
