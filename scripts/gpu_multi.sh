#!/bin/bash
# Multi-GPU session on one box: scripts/gpu_multi.sh <tag> <N> [extra bench flags]
tag=$1; n=$2; shift 2
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_peer_gather_gpu.py -m gpu -x -q > $out/${tag}_pytest_peer.log 2>&1
echo "pytest peer exit $?"; tail -3 $out/${tag}_pytest_peer.log
run() { # name, flags
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port 29533 bench.py --gpus $n "$@" > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
  echo "bench $name exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('$out/${tag}_${name}.json').read().strip().splitlines()[-1])
    print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), d['kernels_ms'], d['gather'], 'verified', d['gather_verified'])
    if d.get('e2e'): print('  e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],1), d['e2e']['link_gbs'], 'frac', round(d['e2e']['link_frac'],3), d['e2e']['cpu_affinity'])
    for k,v in (d.get('configs') or {}).items(): print('  ',k, round(v['ms_per_step'],4), round(v['value']), v.get('decode_ms_max_over_ranks'), v.get('gather_ms_max_over_ranks'), v.get('checksum_gathered_table'))
except Exception as e:
    print('no line', e)
PY
  tail -3 $out/${tag}_${name}.err
}
run peer "$@"
run nccl --nccl-gather --no-e2e --no-configs "$@"
if [ "$UPLOAD_AB" = "1" ]; then   # e2e with the copy-engine upload of the rectangles instead of the fetch kernel
  run e2e_roi --upload roi --no-configs --no-cpu-baseline "$@"
fi
