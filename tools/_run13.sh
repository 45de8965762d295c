set -x
python -m pytest tests -m gpu -q -k "prune or a7 or a8 or a11 or one_shot" 2>&1 | tail -8
python - <<'PY'
import torch, bench, os, json
for flag in ('1','0'):
    os.environ['CPGB_PRUNE_SAMPLED']=flag
    print(flag, json.dumps(bench.prune_table(torch.device('cuda:0'))))
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"sel_|prune_" -c 40 --csv --log-file gpurun_out/r2_prune_ncu.csv python - <<'PY' > /dev/null 2>&1
import torch, bench, os
os.environ['CPGB_PRUNE_SAMPLED']='1'
bench.prune_table(torch.device('cuda:0'), iters=0)
PY
grep -c sel_ gpurun_out/r2_prune_ncu.csv
