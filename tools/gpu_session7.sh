#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest gemm+modules" ; timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_modules.py tests/test_gpu_graphed.py tests/test_gpu_dropout_fusion.py -x -q > $O/r2_s7_pytest.log 2>&1 ; echo "rc=$?" ; tail -5 $O/r2_s7_pytest.log
echo "== r1 timeline" ; (cd _r1 && timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > ../$O/r2_s7_r1_timeline.log 2>&1; cp gpurun_out/timeline_3xtf32_pipe_h256.csv ../$O/r2_s7_timeline_r1.csv)
echo "== head timeline" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_s7_timeline.log 2>&1 ; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_s7_timeline_head.csv
echo "== head timeline h32" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 32 > $O/r2_s7_timeline32.log 2>&1 ; cp $O/timeline_3xtf32_pipe_h32.csv $O/r2_s7_timeline_head_h32.csv
python tools/timeline_summary.py $O/r2_s7_timeline_r1.csv $O/r2_s7_timeline_head.csv $O/r2_s7_timeline_head_h32.csv > $O/r2_s7_timeline_summary.txt 2>&1
echo "== bench head" ; timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s7.json 2> $O/r2_bench_s7.err ; echo "rc=$?"
echo "== bench head ips100" ; timeout 900 python bench.py --steps 150 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s7_nosync.json 2> $O/r2_bench_s7_nosync.err ; echo "rc=$?"
echo "== bench r1" ; (cd _r1 && timeout 600 python bench.py --steps 150 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm > ../$O/r2_s7_r1_bench_nosync.json 2> ../$O/r2_s7_r1_bench.err)
python - <<'PY'
import json
for f in ['r2_bench_s7','r2_bench_s7_nosync','r2_s7_r1_bench_nosync']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], d.get('sync'))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat $O/r2_s7_timeline_summary.txt
