for v in none epi wgrad; do CPGB_DEBUG_SKIP=$v python bench.py --no-extras --steps 30 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'])"; done
python -m pytest tests/test_optim_gpu.py -q 2>&1 | tail -3
