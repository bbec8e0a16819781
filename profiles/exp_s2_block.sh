for tag in "$@"; do
  if [ $tag = default ]; then lib=mixed_graph_admm_b200/_lib/libmga.so; else lib=mixed_graph_admm_b200/_lib/$tag/libmga.so; fi
  echo "== $tag"
  MGA_LIB=$PWD/$lib timeout 300 python profiles/bench_configs.py t288 large20k pems04 --mode streaming 2>&1 | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], round(d['windows_per_s'],1), round(d['frac_of_hbm_peak'],3))"
done
