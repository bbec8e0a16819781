#!/bin/bash
tag=${1:-r02u}
out=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "cluster or tolerance" > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log | cut -c1-300
for v in "" "MGA_CLUSTER_NATURAL=1" "MGA_CLUSTER_TR=3"; do echo "== $v"; env $v timeout 300 python profiles/notebook_call.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['ms_per_solve'],2)) for k,v in d.items()})"; done
