#!/bin/bash
# Strong scaling of one striped map on the GPUs of this box: tools/striped_measure.sh <ngpus> <size> [extra bench args]
# Writes gpurun_out/striped_<size>_<ngpus>gpu.json (one bench.py line) and .err.
n=$1; size=$2; shift 2
out=gpurun_out/striped_${size}_${n}gpu
if [ "$n" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --size $size --steps 5 --warmup 3 --no-cpu-baseline "$@" > $out.json 2> $out.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $n --size $size --steps 5 --warmup 3 --no-cpu-baseline "$@" > $out.json 2> $out.err
fi
echo "== $n GPU(s), ${size}^2: rc=$?"
python - <<P
import json
try:
    d = json.loads(open("$out.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}, d["config"].get("pcg_iterations"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print({k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("no line:", e); print(open("$out.err").read()[-3000:])
P
