#!/bin/bash
# Round-2 evidence session (one B200): tests, headline bench, launch list, ncu --set full captures (exported to
# CSV / text on the box: the .ncu-rep files are too large to bring back), sanitizer logs.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
export_rep() {   # $1 = report stem: raw metrics CSV + details text, then drop the report
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page details > $O/$1.details.txt 2>/dev/null
  rm -f $O/$1.ncu-rep
}
echo "== pytest all" ; timeout 600 python -m pytest tests -m gpu -q --maxfail=20 > $O/r2_final_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -5 $O/r2_final_pytest.log
echo "== bench final" ; timeout 400 python bench.py --steps 150 --warmup 5 > $O/r2_bench_final.json 2> $O/r2_bench_final.err ; echo "rc=$?"
echo "== bench short" ; timeout 300 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_bench_default_short.json 2> $O/r2_bench_default_short.err ; echo "rc=$?"
echo "== bench r1 on this box" ; (cd _r1 && timeout 300 python bench.py --steps 300 --warmup 5 --iter-per-site 1000 --no-cpu-baseline --no-eval-spmm > ../$O/r2_bench_r1_same_box.json 2> ../$O/r2_bench_r1_same_box.err) ; echo "rc=$?"
echo "== timeline" ; timeout 300 python tools/step_timeline.py 3xtf32 pipe 256 > $O/r2_timeline_final.log 2>&1 ; cp $O/timeline_3xtf32_pipe_h256.csv $O/r2_timeline_final_h256.csv; python tools/timeline_summary.py $O/r2_timeline_final_h256.csv > $O/r2_timeline_final_summary.txt 2>&1
echo "== launch list" ; timeout 300 $NCU --metrics gpu__time_duration.sum --profile-from-start off -c 900 --csv --log-file $O/r2_launches_final_steps.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eval-spmm --no-timeline --soak-s 0 --ncu steps > $O/r2_launches_final_steps.log 2>&1 ; echo "rc=$?"
echo "== ncu step kernels" ; timeout 400 $NCU --set full --profile-from-start off -k regex:"gemm_tf32|batch_fill|batch_count|spmm_seg|spmm_csr|ce_fused|ln_act|splitk" -c 30 -o $O/r2_ncu_step_kernels -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eval-spmm --no-timeline --soak-s 0 --ncu steps > $O/r2_ncu_step_kernels.log 2>&1 ; echo "rc=$?"; export_rep r2_ncu_step_kernels
echo "== ncu slice kernels" ; timeout 300 $NCU --set full --profile-from-start off -k regex:slice_multi -c 2 -o $O/r2_ncu_slice -f python bench.py --steps 6 --warmup 3 --iter-per-site 3 --no-cpu-baseline --no-eval-spmm --no-timeline --soak-s 0 --ncu steps > $O/r2_ncu_slice.log 2>&1 ; echo "rc=$?"; export_rep r2_ncu_slice
echo "== ncu fullgraph" ; timeout 400 $NCU --set full --profile-from-start off -k regex:spmm_csr -c 2 -o $O/r2_ncu_spmm_fullgraph -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-timeline --soak-s 0 --ncu fullgraph > $O/r2_ncu_spmm_fullgraph.log 2>&1 ; echo "rc=$?"; export_rep r2_ncu_spmm_fullgraph
echo "== ncu gemm cfg4 3x" ; timeout 300 $NCU --set full -k regex:"gemm_tf32_kernel<128, false, false, 3" -s 3 -c 1 -o $O/r2_ncu_gemm_cfg4_3xtf32 -f python tools/gemm_bench.py "cfg4 mid fwd" > $O/r2_ncu_gemm_cfg4_3xtf32.log 2>&1 ; echo "rc=$?"; export_rep r2_ncu_gemm_cfg4_3xtf32
echo "== ncu gemm cfg4 1x" ; timeout 300 $NCU --set full -k regex:"gemm_tf32_kernel<256, false, false, 1" -s 3 -c 1 -o $O/r2_ncu_gemm_cfg4_tf32 -f python tools/gemm_bench.py "cfg4 mid fwd" > $O/r2_ncu_gemm_cfg4_tf32.log 2>&1 ; echo "rc=$?"; export_rep r2_ncu_gemm_cfg4_tf32
echo "== gemm bench" ; timeout 300 python tools/gemm_bench.py > $O/r2_gemm_bench.jsonl 2> $O/r2_gemm_bench.err ; echo "rc=$?"
echo "== gemm trace" ; timeout 200 python tools/gemm_trace.py > $O/r2_gemm_trace_final.jsonl 2> $O/r2_gemm_trace_final.err ; echo "rc=$?"
echo "== gat bench" ; timeout 300 python bench.py --model gat --n-hidden 512 --n-heads 4 --n-layers 1 --steps 100 --warmup 5 --no-cpu-baseline --no-eval-spmm > $O/r2_bench_gat.json 2> $O/r2_bench_gat.err ; echo "rc=$?"
echo "== gat bench per-head" ; GIST_GAT_BATCH_HEADS=0 timeout 300 python bench.py --model gat --n-hidden 512 --n-heads 4 --n-layers 1 --steps 100 --warmup 5 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2_bench_gat_perhead.json 2> $O/r2_bench_gat_perhead.err ; echo "rc=$?"
echo "== ncu gat" ; timeout 300 $NCU --set full --profile-from-start off -k regex:gat_ -c 8 -o $O/r2_ncu_gat -f python bench.py --model gat --n-hidden 512 --n-heads 4 --n-layers 1 --steps 4 --warmup 3 --no-cpu-baseline --no-eval-spmm --no-timeline --soak-s 0 --ncu steps > $O/r2_ncu_gat.log 2>&1 ; echo "rc=$?"; export_rep r2_ncu_gat
echo "== configs 0/1" ; timeout 300 python tools/bench_gcn_configs.py > $O/r2_bench_gcn_configs.jsonl 2> $O/r2_bench_gcn_configs.err ; echo "rc=$?"
echo "== sanitizer memcheck" ; timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_fused.py tests/test_gpu_batch.py -x -q -k "copy_src_sum or segment_balanced or fused_epilogue or cross_entropy or layer_norm_act or adam or subgraph_bit_exact" > $O/r2_sanitizer_memcheck.log 2>&1 ; echo "rc=$?"; tail -4 $O/r2_sanitizer_memcheck.log
echo "== sanitizer racecheck" ; timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_fused.py -x -q -k "(segment_balanced and 256) or (cross_entropy and 2586) or (masked_ce_loss_and_grad and 2586 and 3xtf32) or colsum" > $O/r2_sanitizer_racecheck.log 2>&1 ; echo "rc=$?"; tail -4 $O/r2_sanitizer_racecheck.log
echo "== sanitizer memcheck gemm+gat" ; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_gat.py -x -q -k "(rowsum and shape1) or (layernorm_epilogue and shape2) or (gat_aggregate and 300) or (batched_heads and dims0)" > $O/r2_sanitizer_memcheck_gemm_gat.log 2>&1 ; echo "rc=$?"; tail -4 $O/r2_sanitizer_memcheck_gemm_gat.log
python - <<'PY'
import json
for f in ['r2_bench_final','r2_bench_default_short','r2_bench_r1_same_box','r2_bench_gat','r2_bench_gat_perhead']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], (d.get('roofline_fullgraph') or {}).get('d602',{}).get('ms_by_variant'), (d.get('roofline_fullgraph') or {}).get('d256',{}).get('ms_by_variant'))
    except Exception as e: print(f,'ERR',e)
PY
du -sh $O; ls -la $O | awk '{print $5, $9}' | sort -n | tail -8
