#!/bin/bash
# e2e (host-buffer C ABI) step time of bench.py against VAR=values (default: T3C_E2E_LANES 1 2 3): tools/e2e_sweep.sh T3C_PIPE_CHUNKS 4 8 16
var=${1:-T3C_E2E_LANES}; shift
for c in ${@:-1 2 3}; do
  env $var=$c python bench.py --steps 12 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$var', $c, 'e2e ms/step', round(d['e2e']['ms_per_step'], 3), 'serial', round(d['e2e']['serial_ms_per_step'], 3), 'Mpix/s', round(d['e2e']['value'], 1))"
done
