for ns in ${NSS:-2 4 6 10}; do
  EBVO_SLICES=$ns python bench.py --steps 5 --warmup 3 --no-cpu-baseline --strong-frames 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slices $ns value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), 'gn ms', round(d['kernels']['gn']['ms_per_step'],2))"
done
