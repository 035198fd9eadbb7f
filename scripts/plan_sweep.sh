# Sweep of the RANSAC wave-plan shape on the bench workload (run on the GPU box):
#   bash scripts/plan_sweep.sh
# first wave = 1/FIRST of the plan, following waves grow by GROWTH; CHUNKS=1 keeps the plan whole.
run() {
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ba 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1:', round(d['value']), 'hyp/s', round(d['ms_per_step'],4), 'ms/step, e2e', round(d['e2e']['value']), d['kernel_ms_per_step'])"
}
PPSFM_RANSAC_CHUNKS=1 run "whole plan"
for cfg in "3 100" "2 100" "4 100" "8 2" "10 3" "5 4"; do
  set -- $cfg
  PPSFM_RANSAC_FIRST=$1 PPSFM_RANSAC_GROWTH=$2 run "first 1/$1, growth x$2"
done
