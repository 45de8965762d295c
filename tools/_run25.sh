python -m pytest tests -m gpu -q -x -k "not ddp_nccl and not two_ranks" 2>&1 | tail -5
for w in spherenet20 resnet50; do python bench.py --workload $w --steps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['ms_per_step'], d['value'], d['loss'], d['regime_task2']['ms_per_step'], d['regime_task2']['loss'])"; done
