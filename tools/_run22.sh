python -m pytest tests -m gpu -q -k "prelu or optim or reference_models" 2>&1 | tail -8
for v in 1 0; do CPGB_FUSE_PRELU=$v python bench.py --workload spherenet20 --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse_prelu=$v', d['ms_per_step'], d['value'], d['loss'], d['regime_task2']['ms_per_step'], d['regime_task2']['loss'])"; done
