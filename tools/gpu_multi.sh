#!/usr/bin/env bash
# development aid: the multi-GPU checks + benches of one iteration on N GPUs (gpurun --gpus N); output in gpurun_out/<tag>_*
#   bash tools/gpu_multi.sh <tag> <N> [workloads...]
set -u
cd "$(dirname "$0")/.."
tag="${1:-mg}"; N="${2:-2}"; shift 2
wls="${*:-cfg3}"
out=gpurun_out
mkdir -p "$out"
nvidia-smi -L | head -n 8
python -m pytest tests/test_multi_gpu.py -m gpu -q -x > "$out/${tag}_pytest_multi.txt" 2>&1
tail -n 4 "$out/${tag}_pytest_multi.txt"
port=29610
for wl in $wls; do
  steps=20; extra=""
  case "$wl" in cfg5) steps=4;; cfg4) steps=8;; esac
  variants="0 $N"
  case "$wl" in cfg5|cfg4) variants="$N";; esac   # the big configs: plain row sharding only (what the planner picks for cfg5 anyway)
  for rs in $variants; do
    [ "$rs" = "0" ] && name="planned" || name="rows"
    port=$((port+1))
    NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus "$N" --workload "$wl" --steps $steps --warmup 3 --row-shards "$rs" > "$out/${tag}_bench_${wl}_n${N}_${name}.json" 2> "$out/${tag}_bench_${wl}_n${N}_${name}.err"
    python - "$out/${tag}_bench_${wl}_n${N}_${name}.json" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    rs=l.get("row_sharded")
    print(sys.argv[1], "| %s | qps %.0f ms %.3f frac %.3f e2e %.0f scan %.3f" % (l["config"]["parallelism"][:34], l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"], l["roofline"]["kernel_ms"]),
          "| rows-sharded:", (("qps %.0f ms %.3f e2e %.0f" % (rs["value"], rs["ms_per_step"], rs["e2e"]["value"])) if rs else None), "|", l.get("parity_full_scan"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1][:-5]+".err").read()[-1500:])
PY
    # when the planner already chose plain row sharding the forced run would repeat it
    python - "$out/${tag}_bench_${wl}_n${N}_${name}.json" "$N" <<'PY' && break
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    sys.exit(0 if l["config"]["parallelism"].startswith(sys.argv[2]+" row shard") else 1)
except Exception:
    sys.exit(1)
PY
  done
done
