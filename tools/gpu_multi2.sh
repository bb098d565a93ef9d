#!/bin/bash
# Round 2, second multi-GPU session (N = 2 by default): the new paths under NCCL — fused step tail with the LayerNorm
# in the GEMM epilogue (hidden / m <= 128), chunked builder, packed segment records, ids staged a step ahead, and the
# row-streaming merge behind the packed all-gather at the ultra-wide width (4096 hidden units per rank).
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== graphed tests (ids staged ahead)"; timeout 200 python -m pytest tests/test_gpu_graphed.py -m gpu -q > $O/r2b_graphed_test.log 2>&1; echo "rc=$?"; tail -2 $O/r2b_graphed_test.log
echo "== nccl wrapper parity test" ; timeout 240 python -m pytest tests/test_gpu_dist_nccl.py -q -rs > $O/r2b_nccl_test_n$N.log 2>&1 ; echo "rc=$?"; tail -4 $O/r2b_nccl_test_n$N.log
echo "== N=1 quick (ids staged ahead vs not)"
timeout 200 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err
GIST_STAGE_IDS_AHEAD=0 timeout 200 python bench.py --steps 150 --warmup 5 --no-cpu-baseline --no-eval-spmm --no-timeline > $O/r2b_bench_n1_nostage.json 2> $O/r2b_bench_n1_nostage.err
echo "== reddit N=$N" ; timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 150 --warmup 5 --no-eval-spmm --no-cpu-baseline > $O/r2b_scale_reddit_n$N.json 2> $O/r2b_scale_reddit_n$N.err ; echo "rc=$?"; tail -2 $O/r2b_scale_reddit_n$N.err
H=$((4096 * N))
echo "== amazon2m hidden $H (4096 per rank) N=$N" ; timeout 420 $TR --master-port 29512 bench.py --gpus $N --shape amazon2m --n-hidden $H --psize 15000 --steps 100 --warmup 5 --no-eval-spmm --no-cpu-baseline > $O/r2b_scale_cfg4_n$N.json 2> $O/r2b_scale_cfg4_n$N.err ; echo "rc=$?"; tail -3 $O/r2b_scale_cfg4_n$N.err
echo "== amazon2m, 4-byte scatter merge (A/B)" ; GIST_MERGE=scatter timeout 420 $TR --master-port 29514 bench.py --gpus $N --shape amazon2m --n-hidden $H --psize 15000 --steps 100 --warmup 5 --no-eval-spmm --no-cpu-baseline --no-timeline > $O/r2b_scale_cfg4_scatter_n$N.json 2> $O/r2b_scale_cfg4_scatter_n$N.err ; echo "rc=$?"
python - <<PY
import json
for f in ['r2b_bench_n1','r2b_bench_n1_nostage','r2b_scale_reddit_n$N','r2b_scale_cfg4_n$N','r2b_scale_cfg4_scatter_n$N']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        s=d.get('sync') or {}
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), d['e2e']['value'], 'sync', {k:s.get(k) for k in ('ms','pack_ms','all_gather_ms','scatter_ms','dispatch_ms','GBps')}, 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'ERR', e)
PY
