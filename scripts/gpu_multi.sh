#!/bin/bash
# N-GPU bench lines (torchrun, one rank per GPU): config 4 (weak scaling) and config 5 (fixed 512-prompt batch: strong)
N=${1:-2}
mkdir -p gpurun_out
for c in 4 5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --config $c --no-cpu > gpurun_out/r2_bench_c${c}_${N}gpu.json 2> gpurun_out/r2_bench_c${c}_${N}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_c${c}_${N}gpu.json").read().strip().splitlines()[-1])
    print("config $c x$N:", d["value"], d["unit"], d["ms_per_step"], "ms", d["scaling"], d["n_gpus"], d.get("sweep_images_per_s"), d["e2e"])
except Exception as e:
    print("config $c x$N FAILED", e, open("gpurun_out/r2_bench_c${c}_${N}gpu.err").read()[-1500:])
PY
done
