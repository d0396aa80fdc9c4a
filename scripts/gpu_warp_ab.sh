#!/bin/bash
# warp-kernel A/B over library variants: scripts/gpu_warp_ab.sh <tag> <variant> [<variant> ...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
csrc=$PWD/mindpose_b200/csrc
for v in "$@"; do
  lib=$csrc/libposecodec_$v.so; [ "$v" = "default" ] && lib=$csrc/libposecodec.so
  export POSECODEC_LIB=$lib
  echo "== $v" | tee -a $out/${tag}_ab.log
  timeout 600 python -m pytest tests -m gpu -x -q -k "warp or affine" > $out/${tag}_${v}_pytest.log 2>&1
  echo "pytest $v exit $?" | tee -a $out/${tag}_ab.log; tail -2 $out/${tag}_${v}_pytest.log | tee -a $out/${tag}_ab.log
  timeout 300 python scripts/kbench.py --iters 20 --only warp 2>&1 | tee -a $out/${tag}_ab.log
done
unset POSECODEC_LIB
if [ -n "$NCU_VARIANT" ]; then
  lib=$csrc/libposecodec_$NCU_VARIANT.so; [ "$NCU_VARIANT" = "default" ] && lib=$csrc/libposecodec.so
  export POSECODEC_LIB=$lib
  CMD="python scripts/kbench.py --iters 3 --only warp"
  timeout 300 $CMD > $out/${tag}_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:warp_affine_u8x3_band -s 3 -c 1 -f -o $out/${tag}_warp $CMD > $out/${tag}_ncu.log 2>&1
  echo "ncu exit $?"
fi
