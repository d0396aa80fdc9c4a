#!/bin/bash
# bottom-up decode iteration: parity tests, kbench line, light ncu counters
tag=${1:-bu}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_bottomup_gpu.py -x -q -k "decode" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -15 $out/${tag}_pytest.log
timeout 240 python scripts/kbench.py --iters 20 --only bottomup > $out/${tag}_kbench.log 2>&1
cat $out/${tag}_kbench.log
timeout 600 ncu --clock-control none -k regex:'bottomup_decode|mask_zero' -s 4 -c 2 \
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio \
  python scripts/kbench.py --iters 3 --only bottomup > $out/${tag}_ncu.log 2>&1
grep -E "bottomup_decode|mask_zero|gpu__time|inst_executed|issue_active|warps_active|dram__|registers" $out/${tag}_ncu.log | head -40
