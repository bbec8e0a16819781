#!/bin/bash
tag=${1:-r02r}
out=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "cluster or tolerance" > $out/${tag}_pytest.log 2>&1; tail -15 $out/${tag}_pytest.log | cut -c1-300
timeout 300 python profiles/notebook_call.py > $out/${tag}_notebook.json 2> $out/${tag}_notebook.err; cat $out/${tag}_notebook.json; tail -3 $out/${tag}_notebook.err
