#!/bin/bash
# one GPU call: the whole -m gpu suite, smoke(), the ncu launch list of one headline step, the same-build ncu --set full
# capture of the dominant qGEMM launch (-> profiles/r2_gemm_traffic.json), the default bench line with its CPU legs and
# the reference arm
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -q -x --no-header 2>&1 | tail -6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== ncu launch list (config 4, one eager step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --config 4 --no-cpu --profile-step > gpurun_out/r2_launches.log 2>&1
wc -l gpurun_out/r2_launches.csv
echo "== ncu --set full on the dominant qGEMM launches (16384x1280x1280, i8 and f16, fp16-out and fp32+resid)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_kernel -c 4 -f -o gpurun_out/r2_gemm_final \
  python scripts/gemm_i8_bench.py --one > gpurun_out/r2_gemm_final.log 2>&1
ncu -i gpurun_out/r2_gemm_final.ncu-rep --page raw --csv \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
  > gpurun_out/r2_gemm_final_dram.csv 2>&1
cat gpurun_out/r2_gemm_final_dram.csv | cut -c1-600
python scripts/make_gemm_traffic.py gpurun_out/r2_gemm_final_dram.csv gpurun_out/r2_gemm_traffic.json
cp gpurun_out/r2_gemm_traffic.json profiles/r2_gemm_traffic.json   # the bench lines below carry this same-build figure
echo "== bench default (config 4, CPU legs)"
timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 600 gpurun_out/r2_bench_default.json
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
tail -c 800 gpurun_out/r2_bench_reference.json
echo "== bench lines of the other BASELINE configs"
for c in 1 2 3 5; do
  timeout 900 python bench.py --config $c --no-cpu > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err
  tail -c 300 gpurun_out/r2_bench_c$c.json; echo
done
echo "== kernel benches"
timeout 200 python scripts/attn_bench.py 2>&1 | tail -7
timeout 300 python scripts/gemm_i8_bench.py > gpurun_out/gemm_i8_bench.log 2>&1; tail -3 gpurun_out/gemm_i8_bench.log | cut -c1-300
