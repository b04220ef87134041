#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
tag="${1:-it4}"
out=gpurun_out; mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -x -q > "$out/${tag}_pytest.txt" 2>&1; tail -n 4 "$out/${tag}_pytest.txt"
timeout 200 python tools/quick_front_bench.py > "$out/${tag}_front_bench.txt" 2>&1; cat "$out/${tag}_front_bench.txt"
timeout 300 python tools/quick_ivf_bench.py > "$out/${tag}_ivf_bench.txt" 2>&1; tail -n 5 "$out/${tag}_ivf_bench.txt"
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"ivf_search|rotate_gemm" -c 12 --csv --log-file "$out/${tag}_launches_ivf.csv" python tools/ncu_probe.py ivf > /dev/null 2>&1
python - "$out/${tag}_launches_ivf.csv" <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1]))); hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; hdr=rows[hi]
kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value"); mu=hdr.index("Metric Unit")
for r in rows[hi+1:]:
    v=float(r[mv].replace(",","")); u=r[mu]; us=v/1000 if u=="ns" else v
    print("  %9.1f us  %s" % (us, r[kn].split("(")[0][-50:]))
PY
