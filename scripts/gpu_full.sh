#!/bin/bash
# One GPU-box session: full parity suite, both bench arms, then ncu (launch list + the two
# headline kernels).  scripts/gpu_full.sh <tag> [ncu]
tag=${1:-run}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 $out/${tag}_pytest_gpu.log
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; python -c "
import json,sys
d=json.loads(open('$out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernels_ms')}, d['roofline']['frac'], d['roofline_other']['frac'], d['e2e'] and (d['e2e']['value'], d['e2e']['link_gbs'], d['e2e']['link_frac']))
for k,v in (d.get('configs') or {}).items(): print(k, v['ms_per_step'], v['value'], [(x['kernel'][:28], round(x['ms'],4), round(x.get('frac',0),3)) for x in v['kernels']])
"; tail -3 $out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
echo "bench ref exit $?"; cut -c1-300 $out/${tag}_bench_reference.json
for sec in decode encode warp bottomup bu_encode refine nms rescale; do
  timeout 240 python scripts/kbench.py --iters 10 --only $sec,group >> $out/${tag}_kbench.log 2>&1 || echo "kbench $sec failed rc=$?" >> $out/${tag}_kbench.log
done
grep -v "exact-pass" $out/${tag}_kbench.log
if [ "$2" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs"
  timeout 300 $CMD > $out/${tag}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k regex:'topdown_|warp_affine|box_to_center|affine_matrices|bottomup_|group_by' -c 60 \
      --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu1.log 2>&1
  echo "ncu launches exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:topdown_decode_kernel -s 3 -c 1 -f -o $out/${tag}_decode $CMD > $out/${tag}_ncu2.log 2>&1
  echo "ncu decode exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:warp_affine_u8x3_band -s 3 -c 1 -f -o $out/${tag}_warp $CMD > $out/${tag}_ncu3.log 2>&1
  echo "ncu warp exit $?"
fi
