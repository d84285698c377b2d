#!/bin/bash
# one GPU call of an epilogue iteration: kernel tests, fused-epilogue timings, i8-vs-f16 GEMM bench (and again with the
# alternate library build _C/libdgq_b200_A.so when present)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_i8_gpu.py tests/test_api_gpu.py tests/test_kernels_gpu.py tests/test_attention_gpu.py -m gpu -q -x --no-header 2>&1 | tail -8
echo "== qkv / geglu epilogues"
timeout 300 python scripts/qkv_bench.py 2>&1 | tail -20
echo "== i8 vs f16"
timeout 300 python scripts/gemm_i8_bench.py > gpurun_out/gemm_i8_bench.log 2>&1
python - <<'PY'
import json
for d in json.load(open("gpurun_out/gemm_i8_bench.json")):
    print(d["m"], d["n"], d["k"], "w", d["wbits"], {k: (v["i8_ms"], v["f16_ms"]) for k, v in d.items() if isinstance(v, dict)})
PY
if [ -f dgq_b200/_C/libdgq_b200_A.so ]; then
  echo "== variant A"
  cp dgq_b200/_C/libdgq_b200_A.so dgq_b200/_C/libdgq_b200.so
  timeout 300 python scripts/gemm_i8_bench.py > gpurun_out/gemm_i8_bench_A.log 2>&1
  python - <<'PY'
import json
for d in json.load(open("gpurun_out/gemm_i8_bench.json")):
    print(d["m"], d["n"], d["k"], "w", d["wbits"], {k: (v["i8_ms"], v["f16_ms"]) for k, v in d.items() if isinstance(v, dict)})
PY
fi
