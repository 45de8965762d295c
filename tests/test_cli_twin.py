"""cpg_b200.cli.cifar100_ddp -- the torchrun twin of CPG_cifar100_main_normal.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for p in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if os.path.isfile(os.path.join(p, 'CPG_cifar100_main_normal.py')):
            return p
    return None


def test_parser_matches_the_reference_command_line():
    """Every option of CPG_cifar100_main_normal.py:30-106 exists in the twin with the same default, type and choices
    (the experiment1 bash loops pass them verbatim)."""
    ref = _ref_root()
    if ref is None:
        pytest.skip('no reference checkout')
    code = ('import sys, json; sys.path.insert(0, %r); sys.argv = ["x"]; import CPG_cifar100_main_normal as m; '
            'print("PARSER" + json.dumps([[a.option_strings, repr(a.default), getattr(a.type, "__name__", None), '
            'list(a.choices) if a.choices else None, type(a).__name__] for a in m.parser._actions]))' % ref)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=ref, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('PARSER')][0]
    want = {tuple(a[0]): a[1:] for a in json.loads(line[len('PARSER'):])}
    from cpg_b200.cli.cifar100_ddp import build_parser
    have = {tuple(a.option_strings): [repr(a.default), getattr(a.type, '__name__', None),
                                      list(a.choices) if a.choices else None, type(a).__name__]
            for a in build_parser()._actions}
    for opts, spec in want.items():
        assert opts in have, opts
        assert have[opts] == spec, (opts, have[opts], spec)
    extra = sorted(o[0] for o in set(have) - set(want))
    assert extra == ['--cpg_root', '--fuse_bn', '--sync_free', '--synthetic'], extra


def _run(args, tmp, nproc=1, timeout=900):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc), '--master-addr',
           '127.0.0.1', '--master-port', '29671', '-m', 'cpg_b200.cli.cifar100_ddp'] + args
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    return subprocess.run(cmd, cwd=str(tmp), capture_output=True, text=True, timeout=timeout, env=env)


def _single(args, tmp, timeout=900):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    return subprocess.run([sys.executable, '-m', 'cpg_b200.cli.cifar100_ddp'] + args, cwd=str(tmp), capture_output=True,
                          text=True, timeout=timeout, env=env)


COMMON = ['--arch', 'custom_vgg_cifar100', '--dataset', 'aquatic_mammals', '--num_classes', '5', '--lr', '1e-2',
          '--lr_mask', '5e-4', '--batch_size', '32', '--weight_decay', '4e-5', '--network_width_multiplier', '1.0',
          '--max_allowed_network_width_multiplier', '1.5', '--total_num_tasks', '20', '--synthetic', '24',
          '--val_batch_size', '64']


@pytest.mark.gpu
def test_twin_exit_codes_checkpoint_and_prune_cycle(tmp_path):
    """Task 1 through the twin on one GPU with synthetic data: the exit-code protocol of experiment1/*.sh, the
    reference checkpoint layout, then a gradual-prune run resumed from that checkpoint and an inference run."""
    if _ref_root() is None or not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref')):
        pytest.skip('baseline/_ref not staged')
    scratch, prune = 'ck/scratch', 'ck/gradual_prune'
    rec = str(tmp_path / 'ck' / 'gradual_prune' / 'record.txt')
    # no accuracy-goal file -> 3 (CPG_cifar100_main_normal.py:175-176)
    r = _single(COMMON + ['--mode', 'finetune', '--save_folder', scratch, '--epochs', '1', '--baseline_acc_file', 'nope.txt'],
                tmp_path)
    assert r.returncode == 3, (r.returncode, r.stderr[-1500:])
    goals = tmp_path / 'goals.txt'
    goals.write_text(json.dumps({'aquatic_mammals': 0.5}))
    # prune mode without a record file -> exit(-1)
    r = _single(COMMON + ['--mode', 'prune', '--save_folder', prune, '--load_folder', scratch, '--epochs', '1',
                          '--baseline_acc_file', str(goals)], tmp_path)
    assert r.returncode == 255, (r.returncode, r.stderr[-1500:])
    # finetune task 1 until it is learnt: exit 0, checkpoint + record written
    r = _single(COMMON + ['--mode', 'finetune', '--save_folder', scratch, '--epochs', '6', '--baseline_acc_file', str(goals),
                          '--pruning_ratio_to_acc_record_file', rec, '--fuse_bn'], tmp_path)
    assert r.returncode == 0, (r.returncode, r.stdout[-1500:], r.stderr[-3000:])
    ck_path = tmp_path / 'ck' / 'scratch' / 'checkpoint-6.pth.tar'
    assert ck_path.is_file(), os.listdir(tmp_path / 'ck' / 'scratch')
    ck = torch.load(str(ck_path), map_location='cpu', weights_only=False)
    assert sorted(ck) == ['dataset2num_classes', 'dataset_history', 'masks', 'model_state_dict', 'shared_layer_info']
    assert ck['dataset_history'] == ['aquatic_mammals'] and len(ck['masks']) == 15
    for name, m in ck['masks'].items():
        assert name.startswith('module.features.') and m.dtype == torch.uint8
        assert bool((m == 1).all())                       # make_finetuning_mask: every free weight now belongs to task 1
    record = json.loads(open(rec).read())
    assert record['0.0'] >= 0.5
    # an impossible goal at a width below the cap -> "expand the network" = 2
    goals.write_text(json.dumps({'aquatic_mammals': 1.5}))
    r = _single(COMMON + ['--mode', 'finetune', '--save_folder', 'ck/scratch_hard', '--epochs', '1', '--baseline_acc_file',
                          str(goals), '--pruning_ratio_to_acc_record_file', str(tmp_path / 'ck' / 'hard' / 'record.txt')],
                tmp_path)
    assert r.returncode == 2, (r.returncode, r.stderr[-1500:])
    goals.write_text(json.dumps({'aquatic_mammals': 0.5}))
    # gradual pruning 0 -> 0.1 resumed from the finetune checkpoint
    r = _single(COMMON + ['--mode', 'prune', '--save_folder', prune, '--load_folder', scratch, '--epochs', '3',
                          '--pruning_interval', '2', '--pruning_frequency', '4', '--initial_sparsity', '0.0',
                          '--target_sparsity', '0.1', '--lr', '1e-3', '--baseline_acc_file', str(goals),
                          '--pruning_ratio_to_acc_record_file', rec], tmp_path)
    assert r.returncode in (0, 6), (r.returncode, r.stderr[-3000:])
    if r.returncode == 0:
        pk = torch.load(str(tmp_path / 'ck' / 'gradual_prune' / '0.1' / 'checkpoint-3.pth.tar'), map_location='cpu',
                        weights_only=False)
        n = sum(m.numel() for m in pk['masks'].values())
        z = sum(int((m == 0).sum()) for m in pk['masks'].values())
        assert abs(z / n - 0.1) < 0.01, z / n             # the cubic schedule reached its target
        assert '0.1' in json.loads(open(rec).read())
        # inference mode re-loads it (load_checkpoint_only_for_evaluate + validate)
        r = _single(COMMON + ['--mode', 'inference', '--load_folder', prune + '/0.1', '--save_folder', prune + '/0.1',
                              '--baseline_acc_file', str(goals)], tmp_path)
        assert r.returncode == 0, (r.returncode, r.stderr[-3000:])


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_twin_two_ranks(tmp_path):
    if not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref')):
        pytest.skip('baseline/_ref not staged')
    goals = tmp_path / 'goals.txt'
    goals.write_text(json.dumps({'aquatic_mammals': 0.5}))
    rec = str(tmp_path / 'ck' / 'gradual_prune' / 'record.txt')
    r = _run(COMMON + ['--mode', 'finetune', '--save_folder', 'ck/scratch', '--epochs', '6', '--baseline_acc_file', str(goals),
                       '--pruning_ratio_to_acc_record_file', rec], tmp_path, nproc=2)
    assert r.returncode == 0, (r.returncode, r.stdout[-1500:], r.stderr[-3000:])
    assert (tmp_path / 'ck' / 'scratch' / 'checkpoint-6.pth.tar').is_file()
