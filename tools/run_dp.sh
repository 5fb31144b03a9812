#!/bin/bash
# One multi-GPU bench line: tools/run_dp.sh N out.json [bench.py flags...]
N=$1; OUT=$2; shift 2
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N "$@" > $OUT 2> ${OUT%.json}.err
tail -2 ${OUT%.json}.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT").read().splitlines() if l.startswith("{")][-1])
    print("$OUT", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["launches_per_step"])
    print("  windows", d["timing"]["window_ms"])
    print("  top", d["top_kernels"][:10])
except Exception as ex:
    print("$OUT", "no JSON line:", ex)
PY
