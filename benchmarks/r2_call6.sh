set -u
mkdir -p gpurun_out
O=gpurun_out
cd benchmarks
run() { echo "== $*"; env "$@" timeout 120 python run_configs.py --only C5 --reps 8 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(round(d['glups'],1), round(d['ms_median'],3), round(d['frac_of_measured_peak'],4))"; }
( run KX_X=1
  run KX_STAR_L2PF=128
  run KX_STAR_L2PF=256
  run KX_X=1
  run KX_STAR_L2PF=128
  run KX_STAR_L2PF=256
) > ../$O/r2_star_sweep6.txt 2>&1
cat ../$O/r2_star_sweep6.txt
