#!/bin/bash
O=gpurun_out/r2final
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tracking.py -m gpu -x -q -k "file or streamed or window or iq" > $O/pytest_file.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_file.log
timeout 900 python bench.py --no-cpu-baseline > $O/bench_track_file.json 2> $O/bench_track_file.err; echo rc=$?
python -c "import json; j=json.loads(open('$O/bench_track_file.json').read().strip().splitlines()[-1]); print(j['value'], j['e2e']['value'], j['e2e']['file_backed'])"
