python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_run30_n2.json 2> gpurun_out/r2_run30_n2.err; tail -3 gpurun_out/r2_run30_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_run30_n2.json').read().strip().splitlines()[-1])
print('N2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'task2', d.get('regime_task2',{}).get('value'), d.get('regime_task2',{}).get('ms_per_step'))
PY
timeout 280 python -m pytest tests/test_ddp_nccl.py tests/test_cli_twin.py -q -m gpu 2>&1 | tail -3
