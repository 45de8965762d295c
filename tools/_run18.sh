set -x
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_run18_tests.log 2>&1; tail -6 gpurun_out/r2_run18_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_run18_n2.json 2> gpurun_out/r2_run18_n2.err; tail -3 gpurun_out/r2_run18_n2.err; cut -c1-300 gpurun_out/r2_run18_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_run18_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('regime_task2',{}).get('ms_per_step'), d.get('regime_task2',{}).get('value'))
PY
