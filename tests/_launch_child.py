"""Child process of tests/test_launch_cpu.py: stands in for cpg_b200.cli.cifar100_ddp under cpg_b200.cli.launch."""
import json
import os
import sys
import time


def main():
    what = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    if what == 'env':                   # record the rendezvous contract, then leave with the given code
        with open(os.path.join(sys.argv[2], 'env%d.json' % rank), 'w') as fh:
            json.dump({k: os.environ.get(k) for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'LOCAL_WORLD_SIZE',
                                                      'MASTER_ADDR', 'MASTER_PORT')}, fh)
        sys.exit(int(sys.argv[3]))
    if what == 'by_rank':               # exit code per rank
        sys.exit(int(sys.argv[2 + rank]))
    if what == 'hang_rank0':            # rank 0 waits "in a collective", the others leave with the given code
        if rank == 0:
            time.sleep(120)
        sys.exit(int(sys.argv[2]))
    if what == 'gloo':                  # the environment is what torch.distributed's env:// rendezvous needs
        import torch
        import torch.distributed as dist
        dist.init_process_group('gloo')
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t)
        ok = float(t) == world * (world + 1) / 2
        dist.destroy_process_group()
        sys.exit(2 if ok else 1)        # 2: the code experiment1's loop reads as "grow the network"
    sys.exit(99)


if __name__ == '__main__':
    main()
