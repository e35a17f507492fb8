for m in 1776; do
PSLAM_SUB_BATCH=$m timeout 600 python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); pk=d['roofline']['per_kernel']; print('$m', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1), 'validate', pk['lsd_validate']['ms_total'], 'regions', pk['lsd_regions']['ms_total'], 'cluster', pk['peac_cluster']['ms_total'], 'flood', pk['peac_flood']['ms_total'])"
done
