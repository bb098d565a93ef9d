#!/bin/bash
# Round 2, session 14: defaults = fused tail + wide-row LN + chunked builder; new: packed segment records + queue
# prefetch in the segment SpMM, shared-memory staging in the row-streaming merge.  Parity, A/B lines, timeline.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm --no-timeline"
echo "== pytest"; timeout 500 python -m pytest tests -m gpu -q --maxfail=30 > $O/r2_s14_pytest.log 2>&1; echo "rc=$?"; tail -4 $O/r2_s14_pytest.log
echo "== pytest old segment path"; GIST_SEG_META=0 GIST_SEG_PREFETCH=0 timeout 300 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_modules.py tests/test_gpu_graphed.py -m gpu -q --maxfail=10 > $O/r2_s14_pytest_oldseg.log 2>&1; echo "rc=$?"; tail -2 $O/r2_s14_pytest_oldseg.log
run() { local name=$1; shift; echo "== bench $name"; env "$@" timeout 240 $B > $O/r2_s14_bench_$name.json 2> $O/r2_s14_bench_$name.err; echo "rc=$?"; }
run default X=1
run noseg GIST_SEG_META=0 GIST_SEG_PREFETCH=0
run meta_only GIST_SEG_META=1 GIST_SEG_PREFETCH=0
run prefetch_only GIST_SEG_META=0 GIST_SEG_PREFETCH=1
run prep2 GIST_PREP_CTAS=2
run prep3 GIST_PREP_CTAS=3
echo "== bench h32 (one rank's share at m = 8)"
env X=1 timeout 240 $B --n-hidden 32 > $O/r2_s14_bench_h32.json 2> $O/r2_s14_bench_h32.err
env GIST_FUSED_TAIL=0 GIST_GEMM_FUSED_LN_WIDE=0 GIST_BUILDER=rows GIST_SEG_META=0 GIST_SEG_PREFETCH=0 timeout 240 $B --n-hidden 32 > $O/r2_s14_bench_h32_old.json 2> $O/r2_s14_bench_h32_old.err
echo "== bench amazon2m h4096"; timeout 300 python bench.py --shape amazon2m --n-hidden 4096 --psize 15000 --steps 60 --warmup 5 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_s14_bench_amz.json 2>$O/r2_s14_bench_amz.err; echo "rc=$?"
echo "== timeline"; timeout 200 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_s14_timeline.log 2>&1; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_s14_timeline.csv 2>/dev/null; python tools/timeline_summary.py $O/r2_s14_timeline.csv > $O/r2_s14_timeline_summary.txt 2>&1
echo "== merge bench"; timeout 240 python tools/merge_bench.py 32768 8 2 100 47 > $O/r2_s14_merge.jsonl 2> $O/r2_s14_merge.err; echo "rc=$?"; cat $O/r2_s14_merge.jsonl
GIST_MERGE_NO_STAGE=1 timeout 240 python tools/merge_bench.py 32768 8 2 100 47 > $O/r2_s14_merge_nostage.jsonl 2> $O/r2_s14_merge_nostage.err; tail -1 $O/r2_s14_merge_nostage.jsonl
python - <<'PY'
import json
for f in ['default','noseg','meta_only','prefetch_only','prep2','prep3','h32','h32_old','amz']:
    try:
        d=json.load(open('gpurun_out/r2_s14_bench_%s.json'%f)); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d.get('gpu_launches'), d['roofline']['achieved'] if 'roofline' in d else None, d['roofline_gemm']['largest'])
    except Exception as e: print(f,'ERR',e)
PY
head -40 $O/r2_s14_timeline_summary.txt
