set -x
python -m pytest tests/test_optim_gpu.py -q -x 2>&1 | tail -25
python -m pytest tests -m gpu -q -k "prune or a7 or a8 or a11 or one_shot" 2>&1 | tail -8
python - <<'PY'
import torch, bench, os, json
for flag in ('1','0'):
    os.environ['CPGB_PRUNE_SAMPLED']=flag
    print(flag, json.dumps(bench.prune_table(torch.device('cuda:0'))))
PY
