#!/bin/bash
# Round 2: compute-sanitizer re-run over the RobotSDF kernels after the instruction diet of robot_serial_kernel.
set -u
OUT=gpurun_out/r02s2
mkdir -p "$OUT"
CS="compute-sanitizer --error-exitcode 7 --launch-timeout 600"
timeout 1500 $CS --tool memcheck python -m pytest tests/test_gpu_composed.py tests/test_gpu_peer.py -m gpu -x -q \
    -k "not test_single_link_robot_contract and not two_ranks" > "$OUT/memcheck_composed_peer.log" 2>&1; echo "rc=$?" >> "$OUT/memcheck_composed_peer.log"
tail -4 "$OUT/memcheck_composed_peer.log"
timeout 1200 $CS --tool racecheck python -m pytest tests/test_gpu_peer.py -m gpu -x -q -k "multi_target" \
    > "$OUT/racecheck_multi_target.log" 2>&1; echo "rc=$?" >> "$OUT/racecheck_multi_target.log"
tail -4 "$OUT/racecheck_multi_target.log"
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" "$OUT"/*.log
