#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest all" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > $O/r2_s8_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -5 $O/r2_s8_pytest.log
for cfg in "bg:" "nobg:GIST_GEMM_BACKGROUND_DW=0" "nors_nobg:GIST_GEMM_BACKGROUND_DW=0 GIST_GEMM_FUSED_ROWSUM=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== bench $name" ; env $envs timeout 900 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_bench_s8_$name.json 2> $O/r2_bench_s8_$name.err ; echo "rc=$?"
done
echo "== bench r1" ; (cd _r1 && timeout 600 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm > ../$O/r2_s8_r1_bench.json 2> ../$O/r2_s8_r1_bench.err)
echo "== head timeline" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_s8_timeline.log 2>&1 ; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_s8_timeline_head.csv
python tools/timeline_summary.py $O/r2_s8_timeline_head.csv > $O/r2_s8_timeline_summary.txt 2>&1
python - <<'PY'
import json
for f in ['r2_bench_s8_bg','r2_bench_s8_nobg','r2_bench_s8_nors_nobg','r2_s8_r1_bench']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
head -30 $O/r2_s8_timeline_summary.txt
