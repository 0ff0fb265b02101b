for sb in ${SBS:-16 20 24 28}; do
  EBVO_SUB_BATCH=$sb python bench.py --steps 5 --warmup 3 --no-cpu-baseline --strong-frames 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SB $sb value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
done
