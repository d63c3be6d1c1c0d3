#!/bin/bash
# Data-parallel variants of bench.py on N GPUs: bucket size, captured (CUDA-graph) DP step, NCCL CTA budget; then BASELINE.json configs[4]
# (DETR encoder) and configs[2] (DeiT-S) data-parallel.  Usage: tools/dp_sweep.sh N OUTPREFIX
N=$1; OUT=$2; PORT=29600
run() {  # name, env...
  name=$1; shift
  PORT=$((PORT+1))
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extras > ${OUT}_${name}.json 2> ${OUT}_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("${OUT}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("${name}", "FAILED", e)
PY
}
run base A=1
run bucket25 VITB200_DP_BUCKET_MB=25
run graph VITB200_DP_GRAPH=1
run graph_bucket25 VITB200_DP_GRAPH=1 VITB200_DP_BUCKET_MB=25
run ctas8 NCCL_MAX_CTAS=8
for cfg in detr_enc deit_s_distill; do
  PORT=$((PORT+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --config $cfg --gpus $N --steps 30 --warmup 5 > ${OUT}_${cfg}.json 2> ${OUT}_${cfg}.err
  tail -c 600 ${OUT}_${cfg}.json | cut -c1-300
done
