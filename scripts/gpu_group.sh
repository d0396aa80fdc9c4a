#!/bin/bash
# A/B of the grouping kernel: parity tests, then timing with the shipped, old and profile builds.
tag=$1
timeout 400 python -m pytest tests -m gpu -x -q -k "group or bottomup_inferencer or max_num" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/${tag}_pytest.log
{
  python scripts/group_prof.py
  for v in gold gprof; do
    [ -f mindpose_b200/csrc/libposecodec_$v.so ] && POSECODEC_LIB=$PWD/mindpose_b200/csrc/libposecodec_$v.so python scripts/group_prof.py
  done
} > gpurun_out/${tag}_group.log 2>&1
cat gpurun_out/${tag}_group.log
