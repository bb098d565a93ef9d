#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest gemm" ; timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_modules.py -x -q > $O/r2_s10_pytest.log 2>&1 ; echo "rc=$?" ; tail -3 $O/r2_s10_pytest.log
echo "== trace"; timeout 300 python tools/gemm_trace.py > $O/r2_s10_trace.jsonl 2>$O/r2_s10_trace.err
echo "== bench"; timeout 600 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s10_bench.json 2>$O/r2_s10_bench.err
echo "== bench amazon"; timeout 600 python bench.py --shape amazon2m --n-hidden 4096 --psize 15000 --steps 60 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s10_bench_amz.json 2>$O/r2_s10_bench_amz.err
echo "== ncu cfg4 gemm"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 6 -c 2 -o $O/r2_ncu_gemm_cfg4 -f python tools/gemm_bench.py "cfg4 mid fwd" > $O/r2_ncu_gemm_cfg4.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ['r2_s10_bench','r2_s10_bench_amz']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['roofline_gemm']['largest'])
    except Exception as e: print(f,'ERR',e)
for l in open('gpurun_out/r2_s10_trace.jsonl'):
    d=json.loads(l)
    if d['variant']=='auto': print('   ',d['shape'],d['tile_n'],d['splits'],'main',d['phase_us_median']['mainloop'],'tot',d['phase_us_median']['cta_total'],'graph',d['graph_us_per_call'])
PY
