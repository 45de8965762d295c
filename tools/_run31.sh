for v in 1 0; do
CPGB_EPILOGUE_STREAM=$v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/ddp_nccl_worker.py > gpurun_out/r2_run31_worker.out 2> gpurun_out/r2_run31_worker.err; echo "epi=$v rc=$?"
grep "AssertionError\|DDP_NCCL_OK" gpurun_out/r2_run31_worker.err gpurun_out/r2_run31_worker.out | head -3 | cut -c1-600
done
