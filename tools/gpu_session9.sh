#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in default OLD_ISSUE TMEM_2BN; do
  if [ $v = default ]; then L=""; else L="GIST_B200_LIB=$PWD/build/ab/libgist_$v.so"; fi
  echo "== trace $v"; env $L timeout 300 python tools/gemm_trace.py > $O/r2_s9_trace_$v.jsonl 2>$O/r2_s9_trace_$v.err
  echo "== bench $v"; env $L timeout 600 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s9_bench_$v.json 2>$O/r2_s9_bench_$v.err
  echo "== bench amazon $v"; env $L timeout 600 python bench.py --shape amazon2m --n-hidden 4096 --psize 15000 --steps 60 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s9_bench_amz_$v.json 2>$O/r2_s9_bench_amz_$v.err
done
echo "== r1 amazon"; (cd _r1 && timeout 600 python bench.py --shape amazon2m --n-hidden 4096 --psize 15000 --steps 60 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm > ../$O/r2_s9_bench_amz_r1.json 2>../$O/r2_s9_bench_amz_r1.err)
python - <<'PY'
import json
for v in ['default','OLD_ISSUE','TMEM_2BN']:
    for f in ['r2_s9_bench_%s'%v,'r2_s9_bench_amz_%s'%v]:
        try:
            d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['roofline_gemm']['largest'])
        except Exception as e: print(f,'ERR',e)
    for l in open('gpurun_out/r2_s9_trace_%s.jsonl'%v):
        d=json.loads(l)
        if d['variant']=='auto' and d['shape'] in ('r3 L1 fwd','r3 dz1','r3 dW0','cfg4 mid fwd','r3 L0 fwd'): print('   ',v,d['shape'],'main',d['phase_us_median']['mainloop'],'tot',d['phase_us_median']['cta_total'],'graph',d['graph_us_per_call'])
try:
    d=json.load(open('gpurun_out/r2_s9_bench_amz_r1.json')); print('r1 amazon', d['value'], d['ms_per_step'], d['roofline_gemm']['largest'])
except Exception as e: print('ERR',e)
PY
