#!/bin/bash
# A/B of experiment libraries on one GPU box (round 2).
#   scripts/gpu_ab.sh <tag>
# For each variant built with `python -m mindpose_b200.csrc.build --variant <name>`:
# parity tests of the affected kernels, then the per-kernel timing table.
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.csv 2>&1
run() {  # name, lib, pytest -k expr, kbench sections
  local name=$1 lib=$2 kexpr=$3 secs=$4
  export POSECODEC_LIB=$lib
  echo "== $name ($lib)" | tee -a $out/${tag}_ab.log
  timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" > $out/${tag}_${name}_pytest.log 2>&1
  echo "pytest $name exit $?" | tee -a $out/${tag}_ab.log
  tail -2 $out/${tag}_${name}_pytest.log | tee -a $out/${tag}_ab.log
  timeout 300 python scripts/kbench.py --iters 20 --only $secs 2>&1 | tee -a $out/${tag}_ab.log
  unset POSECODEC_LIB
}
csrc=$PWD/mindpose_b200/csrc
run base   ""                          "warp or affine or bottomup or decode" warp,bottomup
[ -f $csrc/libposecodec_colmap.so ] && run colmap $csrc/libposecodec_colmap.so "warp or affine" warp
[ -f $csrc/libposecodec_bucoal.so ] && run bucoal $csrc/libposecodec_bucoal.so "bottomup" bottomup
if [ "$2" = "ncu_colmap" ]; then
  export POSECODEC_LIB=$csrc/libposecodec_colmap.so
  CMD="python scripts/kbench.py --iters 3 --only warp"
  timeout 300 $CMD > $out/${tag}_plain_warp.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:warp_affine_u8x3 -s 3 -c 1 -f -o $out/${tag}_warp_colmap $CMD > $out/${tag}_ncu_warp.log 2>&1
  echo "ncu colmap exit $?"
fi
