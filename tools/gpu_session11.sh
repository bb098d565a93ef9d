#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest all" ; timeout 600 python -m pytest tests -m gpu -q --maxfail=20 > $O/r2_s11_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -5 $O/r2_s11_pytest.log
echo "== trace"; timeout 200 python tools/gemm_trace.py "r3" > $O/r2_s11_trace.jsonl 2>$O/r2_s11_trace.err
echo "== bench"; timeout 300 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s11_bench.json 2>$O/r2_s11_bench.err
echo "== bench nostage"; GIST_GEMM_NO_STAGED_EPILOGUE=1 timeout 300 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s11_bench_nostage.json 2>$O/r2_s11_bench_nostage.err
echo "== bench amazon"; timeout 300 python bench.py --shape amazon2m --n-hidden 4096 --psize 15000 --steps 60 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s11_bench_amz.json 2>$O/r2_s11_bench_amz.err
echo "== timeline"; timeout 300 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_s11_timeline.log 2>&1 ; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_s11_timeline_head.csv; python tools/timeline_summary.py $O/r2_s11_timeline_head.csv > $O/r2_s11_timeline_summary.txt 2>&1
python - <<'PY'
import json
for f in ['r2_s11_bench','r2_s11_bench_nostage','r2_s11_bench_amz']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline_gemm']['largest'])
    except Exception as e: print(f,'ERR',e)
for l in open('gpurun_out/r2_s11_trace.jsonl'):
    d=json.loads(l)
    if d['variant']=='auto': print('   ',d['shape'],d['tile_n'],d['splits'],'main',d['phase_us_median']['mainloop'],'tot',d['phase_us_median']['cta_total'],'graph',d['graph_us_per_call'])
PY
head -24 $O/r2_s11_timeline_summary.txt
