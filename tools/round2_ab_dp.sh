#!/bin/bash
# Data-parallel A/B on 2 GPUs (charged 2x):  gpurun --gpus 2 --timeout 1200 -- bash tools/round2_ab_dp.sh
# Question: how much of the 11 % weak-scaling loss (8.08 vs 7.22 ms/step) is the persistent GEMM grids' last CTAs waiting behind
# NCCL's CTAs?  Sweeps the SMs left to NCCL (VLP_DP_RESERVED_SMS -> vlpk_set_reserved_sms) and NCCL's own CTA cap.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
: > gpurun_out/r02_ab_dp.jsonl
run() {
  local name="$1"; shift
  local line
  line=$(env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
         bench.py --gpus 2 --steps 30 --warmup 5 2>gpurun_out/r02_ab_dp_"$name".err | grep '^{' | tail -1)
  echo "{\"ab\": \"$name\", \"line\": ${line:-null}}" >> gpurun_out/r02_ab_dp.jsonl
  echo "$name: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("value"), d.get("ms_per_step"))' 2>/dev/null)"
}
run baseline       VLP_AB=0
run reserve8       VLP_DP_RESERVED_SMS=8
run reserve16      VLP_DP_RESERVED_SMS=16
run reserve24      VLP_DP_RESERVED_SMS=24
run ctas8          NCCL_MAX_CTAS=8
run ctas8_res8     NCCL_MAX_CTAS=8 VLP_DP_RESERVED_SMS=8
run ctas16_res16   NCCL_MAX_CTAS=16 VLP_DP_RESERVED_SMS=16
run groups_1_2_3_3_3 VLP_DP_GROUPS=1,2,3,3,3
run groups_2_2_4_4   VLP_DP_GROUPS=2,2,4,4
