#!/bin/bash
O=gpurun_out/r3x
mkdir -p $O
BDS_TRAJECTORY_JSON=$O/trajectory_1s.json timeout 900 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q -s -k "one_second_trajectory" > $O/pytest_traj.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_traj.log
timeout 900 python tools/trajectory_divergence.py 1000 > $O/trajectory_10s.json 2> $O/trajectory_10s.err; echo "rc=$?"; tail -3 $O/trajectory_10s.err; cat $O/trajectory_10s.json
