#!/bin/bash
# Round-2 GPU session 2: slab SpMM kernel — tests, A/B bench, timeline.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest spmm first" ; timeout 600 python -m pytest tests/test_gpu_spmm.py -x -q > $O/r2_s2_pytest_spmm.log 2>&1 ; echo "rc=$?" ; tail -8 $O/r2_s2_pytest_spmm.log
echo "== pytest all" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > $O/r2_s2_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -30 $O/r2_s2_pytest.log
echo "== bench slab" ; timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline > $O/r2_bench_s2_slab.json 2> $O/r2_bench_s2_slab.err ; echo "rc=$?"; tail -3 $O/r2_bench_s2_slab.err
echo "== bench noslab" ; GIST_SPMM_SLAB=0 timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s2_noslab.json 2> $O/r2_bench_s2_noslab.err ; echo "rc=$?"; tail -3 $O/r2_bench_s2_noslab.err
echo "== bench slab, prep unbalanced" ; GIST_PREP_BALANCED=0 timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s2_slab_prepold.json 2> $O/r2_bench_s2_slab_prepold.err ; echo "rc=$?"
echo "== timeline" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_timeline_s2.log 2>&1 ; echo "rc=$?"; tail -2 $O/r2_timeline_s2.log
python - <<'PY'
import json
for f in ['r2_bench_s2_slab','r2_bench_s2_noslab','r2_bench_s2_slab_prepold']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['by_width'], (d.get('roofline_fullgraph') or {}).get('d602',{}).get('ms_by_variant'), (d.get('roofline_fullgraph') or {}).get('d256',{}).get('ms_by_variant'))
    except Exception as e:
        print(f, 'ERR', e)
PY
