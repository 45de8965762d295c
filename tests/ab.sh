run() { echo "== $1"; env $1 python bench.py --no-extras --steps 40 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value']), round(d['ms_per_step'], 4), round(d['e2e']['value']), d['loss'], d['gpu_launches'])
"; }
run "CPGB_FUSE_BN=1"
run "CPGB_FUSE_BN=0"
