#!/bin/bash
# Round 2, closing evidence (one B200) for the state after the second session's changes (r2c_*): tests, headline bench,
# replay timeline, ncu launch list, ncu --set full of the step's kernels and of the row-streaming merge, sanitizer.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
export_rep() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null; ncu -i $O/$1.ncu-rep --page details > $O/$1.details.txt 2>/dev/null; rm -f $O/$1.ncu-rep; }
echo "== pytest all"; timeout 300 python -m pytest tests -m gpu -q --maxfail=20 > $O/r2c_final_pytest.log 2>&1; echo "rc=$?"; tail -3 $O/r2c_final_pytest.log
echo "== bench final"; timeout 300 python bench.py --steps 150 --warmup 5 > $O/r2c_bench_final.json 2> $O/r2c_bench_final.err; echo "rc=$?"
echo "== timeline"; timeout 120 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2c_timeline.log 2>&1; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2c_timeline_final_h256.csv; python tools/timeline_summary.py $O/r2c_timeline_final_h256.csv > $O/r2c_timeline_final_summary.txt 2>&1
echo "== launch list"; timeout 200 $NCU --metrics gpu__time_duration.sum --profile-from-start off -c 600 --csv --log-file $O/r2c_launches_final_steps.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eval-spmm --no-timeline --soak-s 0 --ncu steps > $O/r2c_launches_final_steps.log 2>&1; echo "rc=$?"
echo "== ncu step kernels"; timeout 240 $NCU --set full --profile-from-start off -k regex:"gemm_tf32|chunk_|spmm_seg|spmm_csr|ce_fused|ln_act|splitk|adam_multi|batch_mark" -c 34 -o $O/r2c_ncu_step_kernels -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eval-spmm --no-timeline --soak-s 0 --ncu steps > $O/r2c_ncu_step_kernels.log 2>&1; echo "rc=$?"; export_rep r2c_ncu_step_kernels
echo "== ncu merge"; timeout 200 $NCU --set full -k regex:"slice_scatter_rows|slice_multi" -c 2 -o $O/r2c_ncu_merge -f python tools/merge_bench.py 32768 8 2 100 47 > $O/r2c_ncu_merge.log 2>&1; echo "rc=$?"; export_rep r2c_ncu_merge
echo "== sanitizer memcheck (new kernels)"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_batch.py tests/test_gpu_fused.py tests/test_gpu_spmm.py -x -q -k "chunked_builder or row_streaming or wrapper_merge or adam_fused or callers_slot or (segment_balanced and 256)" > $O/r2c_sanitizer_memcheck.log 2>&1; echo "rc=$?"; tail -3 $O/r2c_sanitizer_memcheck.log
echo "== sanitizer racecheck (segment kernel with packed records, fused CE, chunk scan)"; timeout 240 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_fused.py tests/test_gpu_batch.py -x -q -k "(segment_balanced and 256) or (masked_ce_loss_and_grad and 2586 and 3xtf32) or (chunked_builder and hubs)" > $O/r2c_sanitizer_racecheck.log 2>&1; echo "rc=$?"; tail -3 $O/r2c_sanitizer_racecheck.log
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c_bench_final.json')); print('final', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
except Exception as e: print('ERR', e)
PY
head -12 $O/r2c_timeline_final_summary.txt
