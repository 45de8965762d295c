for v in 1 0; do CPGB_DGRAD_FIRST=$v python bench.py --no-extras --steps 40 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dgrad_first=$v', d['ms_per_step'], d['loss'])"; done
python tools/sphere_diag.py 60 5e-4 2>&1 | awk 'NR%4==0' | tail -16
