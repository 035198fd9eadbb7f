# sweep of RANSAC knobs on the bench workload: bash scripts/plan_sweep.sh  (GPU box)
for v in 0 1 2 3 4 5 6; do
  PPSFM_SCORE_VARIANT=$v PPSFM_RANSAC_CHUNKS=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ba 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('score variant $v (one wave):', round(d['value']), round(d['ms_per_step'],4), d['kernel_ms_per_step'], d['result'])"
done
