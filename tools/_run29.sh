python bench.py --steps 40 2>/dev/null > gpurun_out/r2_run29_bench.json; python -c "
import json
d=json.loads(open('gpurun_out/r2_run29_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'loss', d['loss'])
print('task2', d['regime_task2']['value'], d['regime_task2']['e2e_value'])
print('prune', d['prune_event']['ms_select_kernels'], 'roofline', d['roofline']['frac'], d['roofline']['by_pass'])"
python -m pytest tests -m gpu -q -x -k "not ddp_nccl and not two_ranks" 2>&1 | tail -3
