timeout 400 python bench.py --shape amazon2m --n-hidden 4096 --no-cpu-baseline --no-eval-spmm --steps 60 > gpurun_out/r1_bench_amazon2m_h4096_final.json 2> gpurun_out/am.err; tail -2 gpurun_out/am.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/r1_bench_amazon2m_h4096_final.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['by_width'],d['roofline_gemm']['achieved'],d['roofline_gemm']['mma_frac'],d['roofline_gemm']['share_of_step'])"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 800 --csv --log-file gpurun_out/r1_launches_amazon2m_final.csv python bench.py --shape amazon2m --n-hidden 4096 --steps 6 --warmup 3 --no-cpu-baseline --no-eval-spmm --ncu steps > gpurun_out/am2.log 2>&1
timeout 300 python bench.py --model gat --no-cpu-baseline --no-eval-spmm > gpurun_out/r1_bench_gat_final.json 2> gpurun_out/gat.err; tail -2 gpurun_out/gat.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/r1_bench_gat_final.json').read().strip().splitlines()[-1]);print(d['metric'],d['value'],d['ms_per_step'],d['e2e']['ms_per_step'])"
