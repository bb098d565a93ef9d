#!/bin/bash
# Round-2 GPU session 1: full GPU test suite, parity spread, GEMM phase trace, default bench, launch list,
# ncu --set full of the step's GEMM kernels.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/r2_s1_gpu.txt 2>&1
echo "== pytest" ; timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_s1_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -5 $O/r2_s1_pytest.log
echo "== parity spread" ; timeout 300 python tools/parity_spread.py --gpu > $O/r2_parity_spread.jsonl 2> $O/r2_parity_spread.err ; echo "rc=$?"; cat $O/r2_parity_spread.jsonl
echo "== gemm trace" ; timeout 300 python tools/gemm_trace.py > $O/r2_gemm_trace.jsonl 2> $O/r2_gemm_trace.err ; echo "rc=$?"; tail -3 $O/r2_gemm_trace.err
echo "== bench" ; timeout 900 python bench.py --steps 150 --warmup 5 > $O/r2_bench_s1.json 2> $O/r2_bench_s1.err ; echo "rc=$?"; tail -3 $O/r2_bench_s1.err; head -c 1500 $O/r2_bench_s1.json
echo "== launch list" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 800 --csv --log-file $O/r2_launches_s1.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eval-spmm --no-timeline --soak-s 0 --ncu steps > $O/r2_launches_s1.log 2>&1 ; echo "rc=$?"
echo "== ncu gemm" ; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tf32 -c 8 -o $O/r2_ncu_gemm_step -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-eval-spmm --no-timeline --soak-s 0 --ncu steps > $O/r2_ncu_gemm_step.log 2>&1 ; echo "rc=$?"
ls -la $O | tail -20
