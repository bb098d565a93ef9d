#!/bin/bash
# Multi-GPU session: usage  gpu_multi.sh N   (N = 2, 4 or 8 GPUs on this box)
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== nccl wrapper parity test" ; timeout 240 python -m pytest tests/test_gpu_dist_nccl.py -q -rs > $O/r2_nccl_test_n$N.log 2>&1 ; echo "rc=$?"; tail -6 $O/r2_nccl_test_n$N.log
echo "== reddit N=$N" ; timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 150 --warmup 5 --no-eval-spmm > $O/r2_scale_reddit_n$N.json 2> $O/r2_scale_reddit_n$N.err ; echo "rc=$?"; tail -3 $O/r2_scale_reddit_n$N.err
H=$((4096 * N))
echo "== amazon2m hidden $H (4096 per rank) N=$N" ; NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 420 $TR --master-port 29512 bench.py --gpus $N --shape amazon2m --n-hidden $H --psize 15000 --steps 100 --warmup 5 --no-eval-spmm > $O/r2_scale_cfg4_n$N.json 2> $O/r2_scale_cfg4_n$N.err ; echo "rc=$?"; grep -v "NCCL INFO" $O/r2_scale_cfg4_n$N.err | tail -5; grep -m3 -i "nvls\|NVLS" $O/r2_scale_cfg4_n$N.err
echo "== reference arm reddit N=$N" ; timeout 240 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/r2_scale_reddit_ref_n$N.json 2> $O/r2_scale_reddit_ref_n$N.err ; echo "rc=$?"
python - <<PY
import json
for f in ['r2_scale_reddit_n$N','r2_scale_cfg4_n$N','r2_scale_reddit_ref_n$N']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), d['e2e']['value'], 'sync', d.get('sync'), 'gemm', (d.get('roofline_gemm') or {}).get('largest'), 'setup', d.get('setup_s'))
    except Exception as e:
        print(f, 'ERR', e)
PY
