#!/bin/bash
# A/B of environment combinations on the headline bench: tools/ab_multi.sh "A=1 B=2" "A=0" ...   (STEPS=200 by default)
steps=${STEPS:-200}
for combo in "$@"; do
  env $combo python bench.py --steps $steps --warmup 5 2>gpurun_out/ab.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-48s %.4f ms  (e2e %.4f ms)  launches/pair %d' % ('$combo', d['ms_per_step'], 1e3/d['e2e']['value'], d['gpu_launches']//d['steps']))
"
done
