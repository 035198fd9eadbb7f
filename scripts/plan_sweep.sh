for cfg in "10 3" "5 4" "5 100" "8 2" "4 3" "20 3" "6 6" "3 100"; do
  set -- $cfg
  PPSFM_RANSAC_FIRST=$1 PPSFM_RANSAC_GROWTH=$2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-ba 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('first/growth $1 $2:', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d['kernel_ms_per_step'], d['gpu_launches'])"
done
