#!/bin/bash
tag=${1:-r02o}
out=gpurun_out
python profiles/e2e_trace.py > $out/${tag}_trace.log 2>&1; tail -6 $out/${tag}_trace.log
python profiles/profile_step.py --mode resident --batch 1024 --steps 5 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench_n1.log 2> $out/${tag}_bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02o_bench_n1.log').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'],l['e2e']['frac_of_value'],'roof',l['roofline']['frac'])
print({k:(v.get('windows_per_s'),v.get('roofline',{}).get('frac')) for k,v in l['probes'].items()})
print(l['cg_iter']['impl']['streaming'])
PY
