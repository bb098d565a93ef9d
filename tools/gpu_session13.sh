#!/bin/bash
# Round 2, session 13: parity of the new paths (fused step tail, wide-row LN in the split-K second pass, builder v2,
# row-streaming merge, programmatic dependent launch) + A/B bench lines for each switch + merge micro-benchmark.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm --no-timeline"
echo "== pytest (everything but the PDL tests)"; timeout 500 python -m pytest tests -m gpu -q --maxfail=30 -k "not programmatic" > $O/r2_s13_pytest.log 2>&1; echo "rc=$?"; tail -4 $O/r2_s13_pytest.log
echo "== pytest PDL"; timeout 200 python -m pytest tests/test_gpu_graphed.py -m gpu -q -k programmatic > $O/r2_s13_pytest_pdl.log 2>&1; echo "rc=$?"; tail -4 $O/r2_s13_pytest_pdl.log
echo "== pytest GIST_PDL=1 (eager launches with the attribute)"; GIST_PDL=1 timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py tests/test_gpu_modules.py tests/test_gpu_dropout_fusion.py tests/test_gpu_spmm.py -m gpu -q --maxfail=10 -k "not programmatic" > $O/r2_s13_pytest_pdl_env.log 2>&1; echo "rc=$?"; tail -3 $O/r2_s13_pytest_pdl_env.log
echo "== pytest GIST_BUILDER=chunks"; GIST_BUILDER=chunks timeout 300 python -m pytest tests/test_gpu_batch.py tests/test_gpu_graphed.py tests/test_gpu_fullsize.py tests/test_gpu_datasets.py -m gpu -q --maxfail=10 -k "not programmatic" > $O/r2_s13_pytest_chunks.log 2>&1; echo "rc=$?"; tail -3 $O/r2_s13_pytest_chunks.log
run() { # name, env...
  local name=$1; shift
  echo "== bench $name"; env "$@" timeout 240 $B > $O/r2_s13_bench_$name.json 2> $O/r2_s13_bench_$name.err; echo "rc=$?"
}
OFF="GIST_FUSED_TAIL=0 GIST_GEMM_FUSED_LN_WIDE=0 GIST_BUILDER=rows GIST_PDL=0"
run base $OFF
run tail GIST_FUSED_TAIL=1 GIST_GEMM_FUSED_LN_WIDE=0 GIST_BUILDER=rows GIST_PDL=0
run tail_ln GIST_FUSED_TAIL=1 GIST_GEMM_FUSED_LN_WIDE=1 GIST_BUILDER=rows GIST_PDL=0
run tail_ln_chunks GIST_FUSED_TAIL=1 GIST_GEMM_FUSED_LN_WIDE=1 GIST_BUILDER=chunks GIST_PDL=0
run all GIST_FUSED_TAIL=1 GIST_GEMM_FUSED_LN_WIDE=1 GIST_BUILDER=chunks GIST_PDL=1
run pdl_only GIST_FUSED_TAIL=0 GIST_GEMM_FUSED_LN_WIDE=0 GIST_BUILDER=rows GIST_PDL=1
echo "== timeline (all on)"; GIST_BUILDER=chunks GIST_PDL=1 timeout 200 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_s13_timeline.log 2>&1; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_s13_timeline_all.csv 2>/dev/null; python tools/timeline_summary.py $O/r2_s13_timeline_all.csv > $O/r2_s13_timeline_all_summary.txt 2>&1
echo "== merge bench (config-4 size, one GPU)"; timeout 240 python tools/merge_bench.py 32768 8 2 100 47 > $O/r2_s13_merge.jsonl 2> $O/r2_s13_merge.err; echo "rc=$?"; cat $O/r2_s13_merge.jsonl
python - <<'PY'
import json
for f in ['base','tail','tail_ln','tail_ln_chunks','all','pdl_only']:
    try:
        d=json.load(open('gpurun_out/r2_s13_bench_%s.json'%f)); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d.get('gpu_launches'), d['roofline_gemm']['largest'])
    except Exception as e: print(f,'ERR',e)
PY
head -30 $O/r2_s13_timeline_all_summary.txt
