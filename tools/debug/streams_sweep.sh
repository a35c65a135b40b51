for s in 2 3 4 6 8; do
  python bench.py --steps 20 --warmup 5 --streams $s --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams', $s, 'ms', round(d['ms_per_step'],4), 'value', round(d['value']/1e9,4), 'e2e', round(d['e2e']['value']/1e9,4))"
done
