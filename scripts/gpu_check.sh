#!/bin/bash
# One GPU-box session: parity tests, per-kernel table, bench line, then (optionally) ncu.
#   scripts/gpu_check.sh [tag] [ncu]
# Everything lands in gpurun_out/<tag>_*.  Each step has its own timeout so that a
# hung kernel cannot hold the box.
tag=${1:-run}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
for sec in decode encode warp bottomup bu_encode refine nms sweep; do
  timeout 240 python scripts/kbench.py --iters 10 --only $sec,group --json $out/${tag}_kbench_$sec.json \
    >> $out/${tag}_kbench.log 2>&1 || echo "kbench $sec failed/timeout rc=$?" >> $out/${tag}_kbench.log
done
cat $out/${tag}_kbench.log
timeout 200 python scripts/e2e_upload_ab.py > $out/${tag}_upload_ab.log 2>&1 || echo "upload A/B failed rc=$?"
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
if [ "$2" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
  timeout 300 $CMD > $out/${tag}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
      -k regex:'topdown_|warp_affine|box_to_center|affine_matrices|bottomup_|group_by' -c 60 \
      --csv --log-file $out/${tag}_launches.csv $CMD > $out/${tag}_ncu1.log 2>&1
  echo "ncu launches exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:topdown_decode_kernel -s 3 -c 1 -f -o $out/${tag}_decode $CMD > $out/${tag}_ncu2.log 2>&1
  echo "ncu decode exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:warp_affine_u8 -s 3 -c 1 -f -o $out/${tag}_warp $CMD > $out/${tag}_ncu3.log 2>&1
  echo "ncu warp exit $?"
  ls -la $out
fi
if [ "$2" = "ncu_bottomup" ]; then
  CMD="python scripts/kbench.py --iters 3 --only bottomup"
  timeout 300 $CMD > $out/${tag}_plain_bu.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:bottomup_decode -s 3 -c 1 -f -o $out/${tag}_bottomup $CMD > $out/${tag}_ncu_bu.log 2>&1
  echo "ncu bottomup exit $?"
fi
