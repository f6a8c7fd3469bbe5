#!/usr/bin/env bash
# prep / fold / DFT ms per step of the small resident bench (16 launches of 1024 chunks per step); extra env via "$@"
cd "$(dirname "$0")/.."
env "$@" python bench.py --chunks 16384 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']; print('$*', 'value', round(d['value']), {k: round(v / 16, 4) for k, v in s.items() if v > 1})"
