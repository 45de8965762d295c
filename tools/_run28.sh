for v in 1 0 1 0; do CPGB_NARROW_FEW_TILES=$v python bench.py --no-extras --steps 40 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('narrow=$v', d['ms_per_step'], d['loss'], d['gpu_launches'])"; done
