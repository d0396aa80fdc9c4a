#!/bin/bash
# warp-kernel iteration: parity tests of the crop warp, then its timing (+ optional ncu)
tag=${1:-warp}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "warp or affine" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -15 $out/${tag}_pytest.log
timeout 300 python scripts/kbench.py --iters 20 --only warp 2>&1 | tee $out/${tag}_kbench.log
if [ "$2" = "ncu" ]; then
  CMD="python scripts/kbench.py --iters 3 --only warp"
  timeout 300 $CMD > $out/${tag}_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:warp_affine_u8x3_band -s 3 -c 1 -f -o $out/${tag}_warp $CMD > $out/${tag}_ncu.log 2>&1
  echo "ncu exit $?"
fi
