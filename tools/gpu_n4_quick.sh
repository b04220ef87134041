#!/usr/bin/env bash
# one planned cfg3 line on 4 GPUs (no tests): gpurun --gpus 4
set -u
cd "$(dirname "$0")/.."
out=gpurun_out; mkdir -p "$out"
NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus 4 --steps 20 --warmup 3 --no-full-parity > "$out/r02c_bench_cfg3_n4_planned.json" 2> "$out/r02c_bench_cfg3_n4_planned.err"
python - "$out/r02c_bench_cfg3_n4_planned.json" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); rs=l.get("row_sharded")
    print("| %s | qps %.0f ms %.3f frac %.3f e2e %.0f scan %.3f" % (l["config"]["parallelism"][:34], l["value"], l["ms_per_step"], l["roofline"]["frac"], l["e2e"]["value"], l["roofline"]["kernel_ms"]),
          "| rows-sharded:", (("qps %.0f ms %.3f e2e %.0f" % (rs["value"], rs["ms_per_step"], rs["e2e"]["value"])) if rs else None))
except Exception as e:
    print("unreadable:", e); print(open(sys.argv[1][:-5]+".err").read()[-1500:])
PY
