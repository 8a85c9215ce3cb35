#!/bin/bash
# First GPU call of round 2 (one B200, ~8 min):  gpurun --timeout 1500 -- bash tools/round2_ab.sh
#  1. the GPU tests written after round 1's budget ran out (skipped by default), each file on its own so one failure does not hide the rest;
#  2. the default bench line, then one line per queued opt-in switch, each with --no-cpu-baseline.
# Results: gpurun_out/r02_unverified_tests.log, gpurun_out/r02_ab.jsonl (one JSON object per configuration, "ab" names it).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
: > gpurun_out/r02_unverified_tests.log
for f in tests/test_zz_abi_split_gpu.py tests/test_zz_bertadam_gpu.py tests/test_zz_fused_head_gpu.py tests/test_zz_table_grads_gpu.py tests/test_zz_tail_split_gpu.py; do
  echo "=== $f" >> gpurun_out/r02_unverified_tests.log
  VLP_RUN_UNVERIFIED=1 timeout 600 python -m pytest "$f" -q -m gpu -p no:cacheprovider >> gpurun_out/r02_unverified_tests.log 2>&1
  echo "exit=$?" >> gpurun_out/r02_unverified_tests.log
done
grep -E "^===|passed|failed|error|exit=" gpurun_out/r02_unverified_tests.log

: > gpurun_out/r02_ab.jsonl
run() {  # name, env assignments...
  local name="$1"; shift
  local line
  line=$(env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02_ab_"$name".err | tail -1)
  echo "{\"ab\": \"$name\", \"line\": ${line:-null}}" >> gpurun_out/r02_ab.jsonl
  echo "$name: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"))' 2>/dev/null)"
}
run baseline          VLP_AB=0
run wgrad_stream      VLPK_WGRAD_STREAM=1
run fused_head        VLP_FUSED_HEAD=1
run fused_tables      VLP_FUSED_TABLE_GRADS=1
run tail_split        VLPK_GEMM_TAIL_SPLIT=1
run mask_pack_warp    VLPK_MASK_PACK_WARP=1
run all               VLPK_WGRAD_STREAM=1 VLP_FUSED_HEAD=1 VLP_FUSED_TABLE_GRADS=1 VLPK_GEMM_TAIL_SPLIT=1 VLPK_MASK_PACK_WARP=1
run baseline_again    VLP_AB=0
