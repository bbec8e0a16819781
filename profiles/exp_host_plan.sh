for plan in default 1024 512,512 296,728 728,296 592,432 432,592 296,296,432 592,296,136 148,296,296,284 296,432,296 200,280,296,248 128,296,296,304; do
  if [ $plan = default ]; then unset MGA_HOST_PLAN; else export MGA_HOST_PLAN=$plan; fi
  python bench.py --no-cpu-baseline --no-cg-probe --steps 20 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); ms = sorted(d['e2e']['ms_per_call']); print('$plan', round(d['e2e']['value']), 'min', ms[0], 'med', ms[len(ms)//2])"
done
