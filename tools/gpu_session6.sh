#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest gemm" ; timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q > $O/r2_s6_pytest_gemm.log 2>&1 ; echo "rc=$?" ; tail -5 $O/r2_s6_pytest_gemm.log
echo "== gemm trace" ; timeout 300 python tools/gemm_trace.py > $O/r2_s6_gemm_trace.jsonl 2> $O/r2_s6_gemm_trace.err ; echo "rc=$?"
echo "== pytest all" ; timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > $O/r2_s6_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -5 $O/r2_s6_pytest.log
for cfg in "fused:" "unfused:GIST_GEMM_FUSED_SPLITK=0 GIST_GEMM_FUSED_ROWSUM=0 GIST_GEMM_FUSED_LN=0" "nors:GIST_GEMM_FUSED_ROWSUM=0" "nosk:GIST_GEMM_FUSED_SPLITK=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  echo "== bench $name" ; env $envs timeout 900 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_s6_$name.json 2> $O/r2_bench_s6_$name.err ; echo "rc=$?"
done
echo "== timeline" ; timeout 600 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_timeline_s6.log 2>&1 ; echo "rc=$?"; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_timeline_s6_h256.csv
python - <<'PY'
import json
for f in ['r2_bench_s6_fused','r2_bench_s6_unfused','r2_bench_s6_nors','r2_bench_s6_nosk']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], d['roofline_gemm']['largest'])
    except Exception as e:
        print(f, 'ERR', e)
for l in open('gpurun_out/r2_s6_gemm_trace.jsonl'):
    d=json.loads(l); print(d['shape'], d['variant'], d['tile_n'], d['splits'], d['kb_per_split'], 'main', d['phase_us_median']['mainloop'], 'tot', d['phase_us_median']['cta_total'], 'graph', d['graph_us_per_call'])
PY
