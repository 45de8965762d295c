for v in 8 16; do CPGB_CLUSTER_SPLITS=$v python bench.py --no-extras --steps 40 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cluster_splits=$v', d['ms_per_step'], d['loss'], d['gpu_launches'])"; done
CPGB_CLUSTER_SPLITS=16 python -m pytest tests -m gpu -q -x -k "full_size or lockstep or golden or bias_grad" 2>&1 | tail -4
python bench.py --workload spherenet20 --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spherenet', d['ms_per_step'], d['value'])"
