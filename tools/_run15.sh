set -x
python -m pytest tests/test_optim_gpu.py -q 2>&1 | tail -5
python bench.py --no-extras --steps 30 > gpurun_out/r2_run15_cpgb.json 2>gpurun_out/r2_run15.err; cut -c1-200 gpurun_out/r2_run15_cpgb.json
CPGB_OPTIM=torch python bench.py --no-extras --steps 30 > gpurun_out/r2_run15_torch.json 2>>gpurun_out/r2_run15.err; cut -c1-200 gpurun_out/r2_run15_torch.json
python - <<'PY'
import json
for n in ('cpgb','torch'):
    d=json.loads(open('gpurun_out/r2_run15_%s.json'%n).read().strip().splitlines()[-1])
    print(n, d['ms_per_step'], d.get('regime_task2',{}).get('ms_per_step'), d['loss'])
PY
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_run15_tests.log 2>&1; tail -5 gpurun_out/r2_run15_tests.log
