"""cpg_b200.cli.launch: the ranks' exit code reaches the bash task loops of experiment1/*.sh (torchrun would fold it
into 1), the rendezvous environment is the one torch.distributed reads, a dead rank does not hang the loop."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(args, timeout=120):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    return subprocess.run([sys.executable, '-m', 'cpg_b200.cli.launch', '--module', 'tests._launch_child'] + args,
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_exit_codes_and_environment(tmp_path):
    r = _launch(['--nproc', '3', '--master_port', '29733', 'env', str(tmp_path), '2'])
    assert r.returncode == 2, (r.returncode, r.stderr[-1000:])
    for rank in range(3):
        e = json.load(open(os.path.join(str(tmp_path), 'env%d.json' % rank)))
        assert e == {'RANK': str(rank), 'LOCAL_RANK': str(rank), 'WORLD_SIZE': '3', 'LOCAL_WORLD_SIZE': '3',
                     'MASTER_ADDR': '127.0.0.1', 'MASTER_PORT': '29733'}
    assert _launch(['--nproc', '2', 'env', str(tmp_path), '0']).returncode == 0
    assert _launch(['--nproc', '2', 'env', str(tmp_path), '6']).returncode == 6


def test_disagreeing_ranks_report_the_first_failure():
    r = _launch(['--nproc', '3', 'by_rank', '0', '5', '3'])
    assert r.returncode == 5 and 'different codes' in r.stderr


def test_a_dead_rank_does_not_hang_the_loop():
    t0 = time.monotonic()
    r = _launch(['--nproc', '2', '--grace', '1', 'hang_rank0', '6'])
    assert r.returncode == 6, (r.returncode, r.stderr[-1000:])     # the code of the rank that left by itself
    assert time.monotonic() - t0 < 60


def test_rendezvous_through_the_launcher():
    r = _launch(['--nproc', '2', '--master_port', '29734', 'gloo'], timeout=300)
    assert r.returncode == 2, (r.returncode, r.stderr[-2000:])


def test_agreed_code_rules():
    from cpg_b200.cli.launch import agreed_code
    assert agreed_code([0, 0]) == 0 and agreed_code([2, 2, 2]) == 2
    assert agreed_code([0, 3, 5]) == 3
    assert agreed_code([-15, 2], killed={0}) == 2 and agreed_code([-15, -15]) == 143
