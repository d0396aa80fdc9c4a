#!/bin/bash
run() { echo "--- cfg=$1"; timeout 60 python scripts/probe_decode.py $1 2>&1 | tail -1; }
run "3855 plain flip noise 20"
run "4096 plain flip noise 20"
run "4096 plain noflip noise 20"
run "8192 orig flip noise 20"
run "8192 shift noflip noise 20"
timeout 300 python tests/stress/stress_decode.py 4096 30 2>&1 | tail -14
