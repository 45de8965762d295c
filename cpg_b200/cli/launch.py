"""One-node, one-process-per-GPU launcher that keeps the exit-code protocol of the reference's bash task loops
("launcher-compatible bash").

``experiment1/CPG_cifar100_scratch_mul_1.5.sh`` drives the 20-task cycle by the exit code of every
``python CPG_cifar100_main_normal.py ...`` call: 2 = "not enough free weights, grow the network and retrain"
(utils/prune.py:41-42), 6 = "accuracy goal missed, stop pruning here", 3 / 5 = configuration errors, 0 = next step.
``torchrun`` folds every non-zero worker exit into its own exit code 1, which breaks those loops.  This launcher starts
the ranks itself (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment, the contract
``torch.distributed`` reads) and exits with the code the ranks agree on -- they always do for the reference's codes:
masks, prune decisions and the all-reduced accuracies are rank-invariant.

In a CPG checkout, one substitution makes the unmodified bash drivers run the twin on every visible GPU::

    sed -i 's/python CPG_cifar100_main_normal.py/python -m cpg_b200.cli.launch --cpg_root ./' experiment1/*.sh
    # and GPU_ID=0,1,2,3,4,5,6,7 at the top of the script (it is exported as CUDA_VISIBLE_DEVICES per call)

Usage: ``python -m cpg_b200.cli.launch [--nproc N] [--master_port P] [--module M] <the reference's flags ...>``.
``--nproc`` defaults to $CPGB_NPROC, else the number of visible CUDA devices.  A rank that dies while the others wait
in a collective does not hang the loop: once one rank has exited with a non-zero code the rest get ``--grace`` seconds
and are then terminated.
"""
import argparse
import os
import signal
import subprocess
import sys
import time


def _visible_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def build_parser():
    p = argparse.ArgumentParser(add_help=False)
    p.add_argument('--nproc', type=int, default=int(os.environ.get('CPGB_NPROC', '0')))
    p.add_argument('--master_addr', type=str, default=os.environ.get('MASTER_ADDR', '127.0.0.1'))
    p.add_argument('--master_port', type=int, default=int(os.environ.get('MASTER_PORT', '29531')))
    p.add_argument('--module', type=str, default='cpg_b200.cli.cifar100_ddp')
    p.add_argument('--grace', type=float, default=60.0)
    return p


def agreed_code(codes, killed=()):
    """Exit code of the job from the ranks' exit codes: the common code if all agree, else the first non-zero one in
    rank order -- among the ranks that exited by themselves when the launcher had to terminate stragglers (`killed`).
    A negative code (killed by that signal) becomes 128 + signal, the shell convention."""
    norm = [(128 - c) if c < 0 else c for c in codes]
    own = [c for r, c in enumerate(norm) if r not in killed] or norm
    if all(c == own[0] for c in own):
        return own[0]
    return next(c for c in own if c != 0)


def launch(nproc, module, rest, master_addr='127.0.0.1', master_port=29531, grace=60.0, env=None):
    base = dict(os.environ if env is None else env)
    procs, killed = [], set()
    for r in range(nproc):
        e = dict(base, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(nproc), LOCAL_WORLD_SIZE=str(nproc),
                 MASTER_ADDR=str(master_addr), MASTER_PORT=str(master_port))
        procs.append(subprocess.Popen([sys.executable, '-m', module] + list(rest), env=e))

    def forward(signum, frame):
        for p in procs:
            if p.poll() is None:
                p.send_signal(signum)
    old = {s: signal.signal(s, forward) for s in (signal.SIGINT, signal.SIGTERM)}
    try:
        deadline = None
        while any(p.poll() is None for p in procs):
            if deadline is None and any(p.poll() not in (None, 0) for p in procs):
                deadline = time.monotonic() + grace          # a rank failed: the others may be stuck in a collective
            if deadline is not None and time.monotonic() > deadline:
                for r, p in enumerate(procs):
                    if p.poll() is None:
                        killed.add(r)
                        p.terminate()
                deadline = float('inf')
            time.sleep(0.05)
    finally:
        for s, h in old.items():
            signal.signal(s, h)
    return [p.returncode for p in procs], killed


def main(argv=None):
    args, rest = build_parser().parse_known_args(argv)
    nproc = args.nproc if args.nproc > 0 else max(_visible_gpus(), 1)
    codes, killed = launch(nproc, args.module, rest, args.master_addr, args.master_port, args.grace)
    code = agreed_code(codes, killed)
    if any(c != codes[0] for c in codes):
        print('cpg_b200.cli.launch: ranks exited with different codes %r -> %d' % (codes, code), file=sys.stderr)
    return code


if __name__ == '__main__':
    sys.exit(main())
