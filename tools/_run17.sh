set -x
python -m pytest tests -m gpu -q -k "fused_bn or padded_channels or optim or bn_model or lockstep or trajectory" 2>&1 | tail -5
python bench.py --no-extras --steps 30 > gpurun_out/r2_run17_a.json 2>gpurun_out/r2_run17.err; cut -c1-200 gpurun_out/r2_run17_a.json
CPGB_BN_CLUSTER=0 python bench.py --no-extras --steps 30 > gpurun_out/r2_run17_b.json 2>>gpurun_out/r2_run17.err; cut -c1-200 gpurun_out/r2_run17_b.json
CPGB_BN_CLUSTER_MB=5 python bench.py --no-extras --steps 30 > gpurun_out/r2_run17_c.json 2>>gpurun_out/r2_run17.err; cut -c1-200 gpurun_out/r2_run17_c.json
