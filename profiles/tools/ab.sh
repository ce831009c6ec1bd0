#!/bin/bash
# A/B helper: prints value (M env-steps/s), ms_per_step, kernel_ms, frac for the current build and env knobs
python bench.py --steps ${STEPS:-1500} --warmup 20 --no-cpu-baseline --no-e2e --no-other-configs 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],4))"
