#!/bin/bash
tag=${1:-bu}
out=gpurun_out
mkdir -p $out
CMD="python scripts/kbench.py --iters 3 --only bottomup"
timeout 300 $CMD > $out/${tag}_plain_bu.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:bottomup_decode -s 3 -c 1 -f -o $out/${tag}_bottomup $CMD > $out/${tag}_ncu_bu.log 2>&1
echo "ncu bottomup exit $?"; tail -3 $out/${tag}_ncu_bu.log
