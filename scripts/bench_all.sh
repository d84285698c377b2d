#!/bin/bash
# bench lines of every BASELINE config (no CPU legs) + the A/B runs of the i8 / implicit-conv / attention-split switches
mkdir -p gpurun_out
for c in 4 2 5 3; do
  timeout 900 python bench.py --config $c --no-cpu > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_c$c.json").read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ("metric", "value", "ms_per_step", "e2e", "launches_per_step", "clocks", "breakdown_ms", "loop_parity", "sweep_images_per_s")}
    keep["roofline"] = {k: d["roofline"].get(k) for k in ("achieved", "frac", "f16", "i8")}
    keep["roofline_attention"] = d.get("roofline_attention", {}).get("frac")
    keep["tail"] = {k: v["frac"] for k, v in d.get("tail", {}).items()}
    print("config $c:", json.dumps(keep))
except Exception as e:
    print("config $c FAILED", e, open("gpurun_out/r2_bench_c$c.err").read()[-1500:])
PY
done
if [ "$1" == "ab" ]; then
for env in "DGQ_I8=0" "DGQ_IMPLICIT_CONV=0" ; do
  echo "== config 5 with $env"; env $env DGQ_LOOP_PARITY=0 timeout 600 python bench.py --config 5 --no-cpu --steps 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['breakdown_ms'])"
done
echo "== config 4 with DGQ_ATTN_SPLIT=0"; DGQ_ATTN_SPLIT=0 timeout 600 python bench.py --config 4 --no-cpu --steps 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['breakdown_ms'])"
fi
