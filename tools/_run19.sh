set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/ddp_nccl_worker.py > gpurun_out/r2_run19_worker.out 2> gpurun_out/r2_run19_worker.err; echo rc=$?
grep "ddp_nccl_worker\|Error\|error" gpurun_out/r2_run19_worker.err | tail -20; tail -3 gpurun_out/r2_run19_worker.out
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -q -k "dataparallel or replica" 2>&1 | tail -5
