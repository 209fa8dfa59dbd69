#!/bin/bash
# Round 2: compute-sanitizer over the kernels written this round (robot_serial / robot_query incl. the cooperative remote
# flush, fk_serial, voxel scatter/gather/compaction, binned-record tree walk + unpermute).
set -u
OUT=gpurun_out/r02s
mkdir -p "$OUT"
CS="compute-sanitizer --error-exitcode 7 --launch-timeout 600"
timeout 1500 $CS --tool memcheck python -m pytest tests/test_gpu_composed.py tests/test_gpu_voxel.py -m gpu -x -q \
    -k "not test_single_link_robot_contract" > "$OUT/memcheck_composed_voxel.log" 2>&1; echo "rc=$?" >> "$OUT/memcheck_composed_voxel.log"
tail -4 "$OUT/memcheck_composed_voxel.log"
timeout 1500 $CS --tool memcheck python -m pytest tests/test_gpu_peer.py tests/test_gpu_mesh.py -m gpu -x -q \
    -k "not large_properties and not two_ranks" > "$OUT/memcheck_peer_mesh.log" 2>&1; echo "rc=$?" >> "$OUT/memcheck_peer_mesh.log"
tail -4 "$OUT/memcheck_peer_mesh.log"
timeout 1200 $CS --tool racecheck python -m pytest tests/test_gpu_peer.py -m gpu -x -q -k "multi_target" \
    > "$OUT/racecheck_multi_target.log" 2>&1; echo "rc=$?" >> "$OUT/racecheck_multi_target.log"
tail -4 "$OUT/racecheck_multi_target.log"
timeout 1200 $CS --tool racecheck python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "edge_cases" \
    > "$OUT/racecheck_voxel.log" 2>&1; echo "rc=$?" >> "$OUT/racecheck_voxel.log"
tail -4 "$OUT/racecheck_voxel.log"
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" "$OUT"/*.log
