python tools/shard_check.py 2>&1 | tail -2
timeout 250 python -m pytest tests/test_ddp_nccl.py -q -m gpu 2>&1 | tail -3
