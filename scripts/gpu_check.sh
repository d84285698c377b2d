#!/bin/bash
# one GPU call: kernel-level tests, the i8-vs-f16 GEMM bench, the attention bench, then the teacher-forced layer tests
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_attention_gpu.py tests/test_fused_gpu.py tests/test_api_gpu.py tests/test_i8_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --no-header 2>&1 | tail -8
timeout 200 python scripts/gemm_i8_bench.py > gpurun_out/gemm_i8_bench.log 2>&1
python - <<'PY'
import json
try:
    for d in json.load(open("gpurun_out/gemm_i8_bench.json")):
        print(d["m"], d["n"], d["k"], "w", d["wbits"], {k: (v["i8_ms"], v["f16_ms"], v["i8_tops"], v["f16_tflops"]) for k, v in d.items() if isinstance(v, dict)})
except Exception as e:
    print("gemm bench:", e); print(open("gpurun_out/gemm_i8_bench.log").read()[-1500:])
PY
timeout 200 python scripts/attn_bench.py 2>&1 | tail -9
if [ "$1" != "nolayer" ]; then
timeout 800 python -m pytest tests/test_layerwise_gpu.py -m gpu -q -s --no-header 2>&1 | grep -E "layerwise\]|passed|failed|^E  " | cut -c1-2300
fi
