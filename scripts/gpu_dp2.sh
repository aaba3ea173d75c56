#!/bin/bash
# 2-GPU train bench under several NCCL CTA caps (NCCL kernels share the SMs with the persistent conv kernels).
set -u
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 30 --warmup 5 --no-subs --no-cpu --no-k1 --no-kernel-profile > gpurun_out/dp2_$tag.json 2> gpurun_out/dp2_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dp2_$tag.json").read().strip().splitlines()[-1]); print("$tag", d["value"], d["ms_per_step"])
except Exception as e: print("$tag failed", e)
PY
}
run default KP_X=1
run min32 NCCL_MIN_CTAS=32
run min64 NCCL_MIN_CTAS=64
