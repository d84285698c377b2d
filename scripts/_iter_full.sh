echo "TWO=0"; DGQ_ATTN_TWO=0 timeout 100 python scripts/attn_bench.py 2>&1 | tail -7 | cut -c20-60 | tr '\n' ' '; echo
for c in 4 5 3; do
DGQ_LOOP_PARITY=0 timeout 600 python bench.py --config $c --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('config $c', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['breakdown_ms'].get('dgq_attention'), d['roofline_attention']['frac'], {k:v for k,v in d['top_shapes_ms'].items() if 'attn' in k})"
done
