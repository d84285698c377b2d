mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fused_gpu.py tests/test_api_gpu.py -m gpu -q -x --no-header 2>&1 | tail -3
for rep in 1 2 3; do timeout 100 python scripts/attn_bench.py 2>&1 | tail -7 | cut -c20-60 | tr '\n' ' '; echo " rc=$?"; done
for c in 4 5 3; do
DGQ_LOOP_PARITY=0 timeout 600 python bench.py --config $c --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('config $c', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['breakdown_ms'].get('dgq_attention'), d['roofline_attention']['frac'], {k:v for k,v in d['top_shapes_ms'].items() if 'attn' in k})"
done
