for cfg in "t288 256 64 32 16 0" "large20k 64 16 8 4 2 0" "pems04 8192 2048 1024 512 0"; do
  set -- $cfg; c=$1; shift
  for G in "$@"; do
    printf "%s group=%s: " $c $G
    MGA_S2_GROUP=$G timeout 300 python profiles/bench_configs.py $c --mode streaming 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['windows_per_s'],1), round(d['frac_of_hbm_peak'],3), d['launches_per_step'], round(d['x_checksum'],4))"
  done
done
