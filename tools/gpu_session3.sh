#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest spmm" ; timeout 600 python -m pytest tests/test_gpu_spmm.py -x -q > $O/r2_s3_pytest_spmm.log 2>&1 ; echo "rc=$?" ; tail -8 $O/r2_s3_pytest_spmm.log
echo "== bench slab" ; timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s3_slab.json 2> $O/r2_bench_s3_slab.err ; echo "rc=$?"; tail -3 $O/r2_bench_s3_slab.err
echo "== bench slab, prep unbalanced" ; GIST_PREP_BALANCED=0 timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s3_slab_prepold.json 2> $O/r2_bench_s3_slab_prepold.err ; echo "rc=$?"
echo "== timeline" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_timeline_s3.log 2>&1 ; echo "rc=$?"; tail -2 $O/r2_timeline_s3.log
python - <<'PY'
import json
for f in ['r2_bench_s3_slab','r2_bench_s3_slab_prepold']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['by_width'])
    except Exception as e:
        print(f, 'ERR', e)
PY
