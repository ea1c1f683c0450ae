#!/bin/bash
# A/B timing of library builds on one box: scripts/ab_kernels.sh libA.so libB.so,ENV=VAL ...
# Each candidate is copied over chrysalis_b200/libchrysalis_b200.so in the box's scratch copy and bench.py is run (twice, interleaved).
for rep in 1 2; do
for arg in "$@"; do
    so=${arg%%,*}; envs=""
    if [[ "$arg" == *,* ]]; then envs=${arg#*,}; envs=${envs//,/ }; fi
    cp "$so" chrysalis_b200/libchrysalis_b200.so
    env $envs python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-pipeline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$arg', 'rep$rep', d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
done
