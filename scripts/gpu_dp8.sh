#!/bin/bash
# 8-GPU train bench: which all-reduce algorithm NCCL picks (NVLS?) and what it costs without it.
set -u
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-subs --no-cpu --no-k1 --no-kernel-profile > gpurun_out/dp8_$tag.json 2> gpurun_out/dp8_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dp8_$tag.json").read().strip().splitlines()[-1]); print("$tag", d["value"], d["ms_per_step"])
except Exception as e: print("$tag failed", e)
PY
}
run info NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING NCCL_DEBUG_FILE=/tmp/nccl_info.%p.log
grep -h -i "nvls\|algo\|channels" /tmp/nccl_info.*.log | sort | uniq -c | sort -rn | head -30 > gpurun_out/dp8_nccl_info.txt
run nonvls NCCL_NVLS_ENABLE=0
