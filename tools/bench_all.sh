#!/bin/bash
# short bench lines of every workload (development aid)
for w in "$@"; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_brief.py; done
