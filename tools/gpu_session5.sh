#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== r1 baseline on this box" ; (cd _r1 && timeout 600 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > ../$O/r2_s5_r1_bench.json 2> ../$O/r2_s5_r1_bench.err) ; echo "rc=$?"
echo "== pytest gemm+fused" ; timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py -x -q > $O/r2_s5_pytest_gemm.log 2>&1 ; echo "rc=$?" ; tail -5 $O/r2_s5_pytest_gemm.log
echo "== gemm trace" ; timeout 300 python tools/gemm_trace.py "r3" > $O/r2_s5_gemm_trace.jsonl 2> $O/r2_s5_gemm_trace.err ; echo "rc=$?"
echo "== bench fused" ; timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline > $O/r2_bench_s5.json 2> $O/r2_bench_s5.err ; echo "rc=$?"; tail -3 $O/r2_bench_s5.err
echo "== bench unfused" ; GIST_GEMM_FUSED_SPLITK=0 GIST_GEMM_FUSED_ROWSUM=0 GIST_GEMM_FUSED_LN=0 timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s5_unfused.json 2> $O/r2_bench_s5_unfused.err ; echo "rc=$?"
echo "== bench fused splitk only" ; GIST_GEMM_FUSED_ROWSUM=0 timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s5_nors.json 2> $O/r2_bench_s5_nors.err ; echo "rc=$?"
echo "== bench rowsum only" ; GIST_GEMM_FUSED_SPLITK=0 timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s5_rsonly.json 2> $O/r2_bench_s5_rsonly.err ; echo "rc=$?"
echo "== timeline" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_timeline_s5.log 2>&1 ; echo "rc=$?"; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_timeline_s5_h256.csv
echo "== timeline unfused" ; GIST_GEMM_FUSED_SPLITK=0 GIST_GEMM_FUSED_ROWSUM=0 GIST_GEMM_FUSED_LN=0 timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_timeline_s5u.log 2>&1 ; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_timeline_s5_h256_unfused.csv
python - <<'PY'
import json
for f in ['r2_s5_r1_bench','r2_bench_s5','r2_bench_s5_unfused','r2_bench_s5_nors','r2_bench_s5_rsonly']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], (d.get('roofline_fullgraph') or {}).get('d602',{}).get('ms_by_variant'), (d.get('roofline_fullgraph') or {}).get('d256',{}).get('ms_by_variant'))
    except Exception as e:
        print(f, 'ERR', e)
PY
