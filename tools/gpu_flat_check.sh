#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
tag="${1:-it6}"
out=gpurun_out; mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -x -q > "$out/${tag}_pytest.txt" 2>&1; tail -n 3 "$out/${tag}_pytest.txt"
for mb in 128 128 64 32 16; do
  echo "== flat chunk $mb MB"
  B200NN_FLAT_CHUNK_MB=$mb timeout 300 python tools/quick_makesearch_bench.py 2>&1 | tail -n 2
done
timeout 120 python tools/quick_flat_bench.py 2>&1 | head -n 1
timeout 200 python tools/quick_ivf_bench.py 2>&1 | tail -n 3
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"flat_|dense_topk|topk_merge|rank_to" -c 40 --csv --log-file "$out/${tag}_launches_flat.csv" python tools/ncu_probe.py flat > /dev/null 2>&1
python - "$out/${tag}_launches_flat.csv" <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1]))); hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; hdr=rows[hi]
kn=hdr.index("Kernel Name"); mv=hdr.index("Metric Value"); mu=hdr.index("Metric Unit")
for r in rows[hi+1:][:16]:
    v=float(r[mv].replace(",","")); u=r[mu]; us=v/1000 if u=="ns" else v
    print("  %9.1f us  %s" % (us, r[kn].split("(")[0][-50:]))
PY
