#!/bin/bash
# development loop: GPU tests (quiet) + a short device-resident bench; prints one summary line
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/dev.json 2> gpurun_out/dev.err || tail -5 gpurun_out/dev.err
python tools/quick_summary.py gpurun_out/dev.json
