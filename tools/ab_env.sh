#!/bin/bash
# A/B of one environment switch on the headline bench: tools/ab_env.sh VAR "v1 v2 ..." [repeats]
var=$1; vals=$2; rep=${3:-2}
for r in $(seq $rep); do for v in $vals; do env $var=$v python bench.py --steps 100 --warmup 5 2>gpurun_out/ab.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$var=$v', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), ' launches', d['gpu_launches'], ' d_loss', d['detail']['final_d_loss'])
"; done; done
