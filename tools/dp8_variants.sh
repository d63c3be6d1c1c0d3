#!/bin/bash
# 8-GPU NCCL variants of bench.py on ONE box, plus the 1-GPU line of the same box (same-box scaling efficiency).  Usage: tools/dp8_variants.sh OUTPREFIX
OUT=$1; PORT=29800
run() {  # name, env...
  name=$1; shift
  PORT=$((PORT+1))
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-extras --no-roofline > ${OUT}_${name}.json 2> ${OUT}_${name}.err
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
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --no-roofline > ${OUT}_1gpu.json 2> ${OUT}_1gpu.err
python - <<PY
import json
d = json.loads(open("${OUT}_1gpu.json").read().strip().splitlines()[-1])
print("1gpu", "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "clk", d["clocks"]["sm_mhz"])
PY
