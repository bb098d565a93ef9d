#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest subset" ; timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_modules.py tests/test_gpu_graphed.py tests/test_gpu_dropout_fusion.py -x -q > $O/r2_s12_pytest.log 2>&1 ; echo "rc=$?" ; tail -3 $O/r2_s12_pytest.log
for cfg in "dzfirst_bg:" "dzfirst_nobg:GIST_GEMM_BACKGROUND_DW=0" "old_bg:GIST_DZ_FIRST=0" "old_nobg:GIST_DZ_FIRST=0 GIST_GEMM_BACKGROUND_DW=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== bench $name" ; env $envs timeout 300 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s12_bench_$name.json 2> $O/r2_s12_bench_$name.err ; echo "rc=$?"
done
echo "== timeline"; timeout 300 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_s12_timeline.log 2>&1 ; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_s12_timeline_head.csv; python tools/timeline_summary.py $O/r2_s12_timeline_head.csv > $O/r2_s12_timeline_summary.txt 2>&1
python - <<'PY'
import json
for f in ['dzfirst_bg','dzfirst_nobg','old_bg','old_nobg']:
    try:
        d=json.load(open('gpurun_out/r2_s12_bench_%s.json'%f)); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
head -16 $O/r2_s12_timeline_summary.txt
