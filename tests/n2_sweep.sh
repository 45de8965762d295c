run() { echo "== $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value']), round(d['ms_per_step'], 4), round(d['e2e']['value']))
"; }
run "CPGB_SM_MARGIN=16" 29521
run "NCCL_MAX_CTAS=24 CPGB_SM_MARGIN=24" 29522
run "NCCL_MAX_CTAS=32 CPGB_SM_MARGIN=32" 29523
run "NCCL_MAX_CTAS=16 CPGB_SM_MARGIN=8" 29524
run "NCCL_MAX_CTAS=12 CPGB_SM_MARGIN=12" 29525
