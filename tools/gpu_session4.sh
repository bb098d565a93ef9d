#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest gemm" ; timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q > $O/r2_s4_pytest_gemm.log 2>&1 ; echo "rc=$?" ; tail -25 $O/r2_s4_pytest_gemm.log
echo "== pytest all" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > $O/r2_s4_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -25 $O/r2_s4_pytest.log
echo "== bench fused" ; timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline > $O/r2_bench_s4.json 2> $O/r2_bench_s4.err ; echo "rc=$?"; tail -3 $O/r2_bench_s4.err
echo "== bench unfused" ; GIST_GEMM_FUSED_SPLITK=0 GIST_GEMM_FUSED_ROWSUM=0 GIST_GEMM_FUSED_LN=0 timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s4_unfused.json 2> $O/r2_bench_s4_unfused.err ; echo "rc=$?"
echo "== timeline" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_timeline_s4.log 2>&1 ; echo "rc=$?"; tail -2 $O/r2_timeline_s4.log; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_timeline_s4_h256.csv
echo "== timeline m8 width" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 32 > $O/r2_timeline_s4_h32.log 2>&1 ; echo "rc=$?"; cp $O/timeline_3xtf32_pipe_h32.csv $O/r2_timeline_s4_h32.csv
python - <<'PY'
import json
for f in ['r2_bench_s4','r2_bench_s4_unfused']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], d['roofline']['by_width'], (d.get('roofline_fullgraph') or {}).get('d602',{}).get('ms_by_variant'), (d.get('roofline_fullgraph') or {}).get('d256',{}).get('ms_by_variant'), d['replay_timeline']['kernels_per_step'] if d.get('replay_timeline') else None)
    except Exception as e:
        print(f, 'ERR', e)
PY
