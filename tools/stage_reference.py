#!/usr/bin/env python
"""Stage the UNMODIFIED reference checkout (ivclab/CPG) under baseline/_ref/.

    python tools/stage_reference.py [/root/reference]

baseline/_ref/ is git-ignored (the reference's sources never enter this repository's history) but is NOT
gpurun-ignored, so the copy travels with the snapshot to the GPU box, where /root/reference does not exist.
Two things run from it there, both through the reference's own public API:

  * ``bench.py --impl reference``: the reference's ``models.custom_vgg_cifar100`` + ``utils.prune.SparsePruner``
    on the host cores (the CPU arm of the headline ratio);
  * ``tests/test_reference_manager_gpu.py``: the reference's ``utils.manager.Manager.train`` /
    ``save_checkpoint`` / ``load_checkpoint`` driving the cpg_b200 layers after ``cpg_b200.install()``.

Only Python sources and the experiment shell scripts are copied (docs/ holds 11 MB of slides).  Nothing is
edited; a MANIFEST with the sha256 of every staged file is written next to them.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, 'baseline', '_ref')
KEEP_DIRS = ('models', 'utils', 'tools', 'experiment1', 'experiment2', 'experiment3', 'packnet_models')


def stage(src='/root/reference'):
    if not os.path.isdir(os.path.join(src, 'models')):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for name in sorted(os.listdir(src)):
        p = os.path.join(src, name)
        if os.path.isdir(p) and name in KEEP_DIRS:
            shutil.copytree(p, os.path.join(DST, name), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        elif os.path.isfile(p) and (name.endswith('.py') or name in ('LICENSE', 'README.md')):
            shutil.copy2(p, os.path.join(DST, name))
    lines = []
    for base, _, files in sorted(os.walk(DST)):
        for f in sorted(files):
            q = os.path.join(base, f)
            lines.append('%s  %s' % (hashlib.sha256(open(q, 'rb').read()).hexdigest(), os.path.relpath(q, DST)))
    with open(os.path.join(DST, 'MANIFEST.sha256'), 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    return True


if __name__ == '__main__':
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print('staged baseline/_ref' if ok else 'reference checkout not found: nothing staged')
