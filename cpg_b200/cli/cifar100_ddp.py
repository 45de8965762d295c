#!/usr/bin/env python
"""torchrun / DDP twin of the reference's ``CPG_cifar100_main_normal.py``.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m cpg_b200.cli.cifar100_ddp \\
        --cpg_root /path/to/CPG --arch custom_vgg_cifar100 --dataset aquatic_mammals --num_classes 5 ... (same flags)

Same command line (CPG_cifar100_main_normal.py:30-106), same exit-code protocol for the bash task loops of
experiment1/*.sh (2 = grow the network, 3 = no accuracy goal file, 5 = no free space, 6 = stop pruning, 255 = prune
without a record file, 1 = bad arch / num_classes), same checkpoint dictionary -- written by rank 0 only, mask keys
keep the ``module.`` prefix -- and the same per-epoch learning-rate schedule.  What changes is the parallelism: the
reference wraps the model in single-process ``nn.DataParallel`` (:199); here every rank owns one GPU, runs the
UNMODIFIED ``utils.manager.Manager`` on its shard of each batch (``--batch_size`` stays the GLOBAL batch), and the
gradients meet in cpg_b200.ddp.GradAllReducer, which the product SparsePruner calls at the top of
``do_weight_decay_and_make_grads_zero`` -- the one call ``Manager.train`` makes between ``backward()`` and
``optimizers.step()`` (utils/manager.py:64-70).  Validation runs on every rank over the whole validation set:
``apply_mask`` is destructive (utils/prune.py:223-231) and must happen on all replicas, and identical metrics give
identical exit codes.  Training accuracy (computed per shard) is averaged across ranks before it decides anything.

``--synthetic N`` replaces the ImageFolder loaders by N seeded random batches per epoch (no dataset on this box).
"""
import argparse
import json
import logging
import math
import os
import sys

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

VGG16_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']


def build_parser():
    """The reference's options, name for name and default for default (CPG_cifar100_main_normal.py:30-106)."""
    p = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    a = p.add_argument
    a('--arch', type=str, default='resnet50', help='Architectures')
    a('--num_classes', type=int, default=-1, help='Num outputs for dataset')
    a('--lr', type=float, default=0.1, help='Learning rate for parameters, used for baselines')
    a('--lr_mask', type=float, default=1e-4, help='Learning rate for mask')
    a('--lr_mask_decay_every', type=int, help='Step decay every this many epochs')
    a('--batch_size', type=int, default=32, help='input batch size for training (global, split across ranks)')
    a('--val_batch_size', type=int, default=100, help='input batch size for validation')
    a('--workers', type=int, default=24, help='')
    a('--weight_decay', type=float, default=0.0, help='Weight decay')
    a('--mask_init', default='1s', choices=['1s', 'uniform', 'weight_based_1s'], help='Type of mask init')
    a('--mask_scale', type=float, default=1e-2, help='Mask initialization scaling')
    a('--mask_scale_gradients', type=str, default='none', choices=['none', 'average', 'individual'],
      help='Scale mask gradients by weights')
    a('--threshold_fn', choices=['binarizer', 'ternarizer'], help='Type of thresholding function')
    a('--threshold', type=float, default=2e-3, help='')
    a('--dataset', type=str, default='', help='Name of dataset')
    a('--train_path', type=str, default='', help='Location of train data')
    a('--val_path', type=str, default='', help='Location of test data')
    a('--save_prefix', type=str, default='checkpoints/', help='Location to save model')
    a('--cuda', action='store_true', default=True, help='use CUDA')
    a('--seed', type=int, default=1, help='random seed')
    a('--checkpoint_format', type=str, default='./{save_folder}/checkpoint-{epoch}.pth.tar',
      help='checkpoint file format')
    a('--epochs', type=int, default=160, help='number of epochs to train')
    a('--restore_epoch', type=int, default=0, help='')
    a('--image_size', type=int, default=32, help='')
    a('--save_folder', type=str, help='folder name inside one_check folder')
    a('--load_folder', default='', help='')
    a('--pruning_interval', type=int, default=100, help='')
    a('--pruning_frequency', type=int, default=10, help='')
    a('--initial_sparsity', type=float, default=0.0, help='')
    a('--target_sparsity', type=float, default=0.1, help='')
    a('--mode', choices=['finetune', 'prune', 'inference'], help='Run mode')
    a('--baseline_acc_file', type=str, help='file to restore baseline validation accuracy')
    a('--network_width_multiplier', type=float, default=1.0, help='the multiplier to scale up the channel width')
    a('--test_piggymask', action='store_true', default=False, help='')
    a('--pruning_ratio_to_acc_record_file', type=str, help='')
    a('--allow_acc_diff', type=float, help='')
    a('--finetune_again', action='store_true', default=False, help='')
    a('--max_allowed_network_width_multiplier', type=float, help='')
    a('--log_path', type=str, help='')
    a('--total_num_tasks', type=int, help='')
    # additions of the twin (absent from the reference command line)
    a('--cpg_root', type=str, default=os.environ.get('CPG_ROOT', ''),
      help='CPG checkout providing models/ and utils/ (default: $CPG_ROOT, else baseline/_ref of this repository)')
    a('--synthetic', type=int, default=0, help='N > 0: N seeded random batches per epoch instead of the ImageFolder data')
    a('--fuse_bn', action='store_true', default=False,
      help='swap BatchNorm2d+ReLU(+MaxPool2d) for the cpg_b200.fused_norm kernels (state_dict keys unchanged)')
    a('--sync_free', action='store_true', default=False,
      help='run the epochs through cpg_b200.train_loop.train_sync_free (Manager.train without its per-batch host reads)')
    return p


class RankModule(nn.Module):
    """What nn.DataParallel is to the reference (CPG_cifar100_main_normal.py:199) without its replication: exposes
    ``.module`` and therefore the ``module.``-prefixed parameter / mask names of the reference checkpoints."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


class ShardedBatches:
    """Iterates a global-batch loader and hands this rank its contiguous shard (SURVEY 8e 'Partitioning')."""

    def __init__(self, loader, rank, world):
        self.loader, self.rank, self.world = loader, rank, world

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for data, target in self.loader:
            n = data.shape[0] // self.world
            if n == 0:
                continue
            yield data[self.rank * n:(self.rank + 1) * n], target[self.rank * n:(self.rank + 1) * n]


def synthetic_loader(n_batches, batch, classes, seed):
    """Seeded, learnable stand-in for a CIFAR-100 superclass: every class is a fixed random 3x32x32 pattern plus noise
    (the patterns depend on `classes` only, so the train and validation loaders agree)."""
    patterns = torch.randn(classes, 3, 32, 32, generator=torch.Generator().manual_seed(classes))
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        target = torch.randint(0, classes, (batch,), generator=g)
        out.append((patterns[target] + 0.5 * torch.randn(batch, 3, 32, 32, generator=g), target))
    return out


def _dist_env():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    return world, int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))


def _all_mean(value, world, device):
    if world <= 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return float(t.item()) / world


def _latest_epoch(fmt, folder):
    """CPG_cifar100_main_normal.py:141-148: the newest checkpoint-<epoch> file in `folder`, 0 if none."""
    for epoch in range(200, 0, -1):
        if os.path.exists(fmt.format(save_folder=folder, epoch=epoch)):
            return epoch
    return 0


def _sharable(model, nl):
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (nl.SharableConv2d, nl.SharableLinear))]


def _fresh_mask(module, device):
    return torch.zeros(module.weight.shape, dtype=torch.uint8, device=device)


def _fit_masks(model, masks, nl, mode, device):
    """:209-254 -- after the network was grown (or an older, narrower task is evaluated) the stored task masks are
    copied into / cut out of masks of the current shapes."""
    grow = shrink = False
    for name, m in _sharable(model, nl):
        if m.weight.dim() == 4:
            if masks[name].size(1) < m.weight.size(1):
                assert mode == 'finetune'
                grow = True
            elif masks[name].size(1) > m.weight.size(1):
                assert mode == 'inference'
                shrink = True
    if not (grow or shrink):
        return
    for name, m in _sharable(model, nl):
        new, old = _fresh_mask(m, device), masks[name]
        if grow and mode == 'finetune':
            new[tuple(slice(0, s) for s in old.shape[:2])].copy_(old)
        elif mode == 'inference':
            new.copy_(old[tuple(slice(0, s) for s in new.shape[:2])])
        else:
            continue
        masks[name] = new


def _new_piggymasks(model, masks, nl):
    """:262-270: one real-valued mask per sharable layer, every element 0.01."""
    for name, m in _sharable(model.module, nl):
        m.piggymask = Parameter(torch.full_like(masks['module.' + name], 0.01, dtype=torch.float32))


def main(argv=None):
    args = build_parser().parse_args(argv)
    world, rank, local = _dist_env()
    # :115-116 the bash loops pass the area multiplier; channel counts scale with its square root
    args.network_width_multiplier = math.sqrt(args.network_width_multiplier)
    args.max_allowed_network_width_multiplier = math.sqrt(args.max_allowed_network_width_multiplier)
    if args.mode == 'prune':
        args.save_folder = os.path.join(args.save_folder, str(args.target_sparsity))
        if args.initial_sparsity != 0.0:
            args.load_folder = os.path.join(args.load_folder, str(args.initial_sparsity))
    if rank == 0:
        if args.save_folder and not os.path.isdir(args.save_folder):
            os.makedirs(args.save_folder, exist_ok=True)
        rec = args.pruning_ratio_to_acc_record_file
        if rec and os.path.dirname(rec) and not os.path.isdir(os.path.dirname(rec)):
            os.makedirs(os.path.dirname(rec), exist_ok=True)
    if rank != 0:
        os.environ.setdefault('TQDM_DISABLE', '1')

    here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    root = args.cpg_root or os.path.join(here, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(root, 'models')):
        print('no CPG checkout at %r (pass --cpg_root)' % root)
        sys.exit(1)
    sys.path.insert(0, root)
    import cpg_b200
    from cpg_b200 import ddp
    nl, _ = cpg_b200.install()
    import models
    import utils
    from utils import Optimizers, set_logger
    from utils.manager import Manager

    if args.log_path and rank == 0:
        set_logger(args.log_path)
    if not torch.cuda.is_available():
        logging.info('no gpu device available')
        print('cpg_b200 has no CPU path: a CUDA device is required')
        sys.exit(1)
    if world > 1:
        import torch.distributed as dist
        ddp.tune_env(world)
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    cpg_b200.torch_compat_shims(device)

    resume_folder = args.load_folder
    resume_from_epoch = args.restore_epoch or _latest_epoch(args.checkpoint_format, resume_folder)
    utils.set_dataset_paths(args)
    if resume_from_epoch:
        ck = torch.load(args.checkpoint_format.format(save_folder=resume_folder, epoch=resume_from_epoch),
                        map_location=device, weights_only=False)
        dataset_history, dataset2num_classes = ck['dataset_history'], ck['dataset2num_classes']
        masks, shared_layer_info = ck['masks'], ck['shared_layer_info']
        if args.mode == 'inference' and 'network_width_multiplier' in shared_layer_info[args.dataset]:
            args.network_width_multiplier = shared_layer_info[args.dataset]['network_width_multiplier']
    else:
        dataset_history, dataset2num_classes, masks, shared_layer_info = [], {}, {}, {}

    if args.baseline_acc_file is None or not os.path.isfile(args.baseline_acc_file):
        sys.exit(3)
    with open(args.baseline_acc_file, 'r') as fh:
        baseline_acc = float(json.load(fh)[args.dataset])
    if args.mode == 'prune' and not args.pruning_ratio_to_acc_record_file:
        sys.exit(-1)

    kw = dict(dataset_history=dataset_history, dataset2num_classes=dataset2num_classes,
              network_width_multiplier=args.network_width_multiplier, shared_layer_info=shared_layer_info)
    if args.arch == 'resnet18':
        model = models.__dict__[args.arch](**kw)
    elif 'vgg' in args.arch:
        model = models.__dict__[args.arch](list(VGG16_CFG), **kw)
    else:
        print('Error!')
        sys.exit(1)
    model.add_dataset(args.dataset, args.num_classes)
    model.set_dataset(args.dataset)
    if args.fuse_bn:
        from cpg_b200.fused_norm import fuse_bn_relu
        fuse_bn_relu(model)
    model = RankModule(model).to(device)

    if not masks:
        for name, m in _sharable(model, nl):
            masks[name] = _fresh_mask(m, device)
    else:
        for name in list(masks):
            masks[name] = masks[name].to(device)
        _fit_masks(model, masks, nl, args.mode, device)

    task_id = model.module.datasets.index(args.dataset) + 1
    if args.dataset not in shared_layer_info:
        shared_layer_info[args.dataset] = {k: {} for k in ('bias', 'bn_layer_running_mean', 'bn_layer_running_var',
                                                            'bn_layer_weight', 'bn_layer_bias', 'piggymask')}
        if task_id > 1:
            _new_piggymasks(model, masks, nl)
    elif args.finetune_again:
        _new_piggymasks(model, masks, nl)
    elif task_id > 1:
        stored = shared_layer_info[args.dataset]['piggymask']
        for name, m in _sharable(model.module, nl):
            m.piggymask = stored[name]
    shared_layer_info[args.dataset]['network_width_multiplier'] = args.network_width_multiplier

    if args.num_classes not in (2, 5):
        print('num_classes should be either 2 or 5')
        sys.exit(1)
    if args.synthetic > 0:
        train_loader = synthetic_loader(args.synthetic, args.batch_size, args.num_classes, args.seed + 100)
        val_loader = synthetic_loader(max(1, args.synthetic // 4), args.val_batch_size, args.num_classes, args.seed + 200)
    else:
        import utils.cifar100_dataset as dataset
        two = args.num_classes == 2
        train_loader = (dataset.cifar100_train_loader_two_class if two else dataset.cifar100_train_loader)(
            args.dataset, args.batch_size)
        val_loader = (dataset.cifar100_val_loader_two_class if two else dataset.cifar100_val_loader)(
            args.dataset, args.val_batch_size)
    if world > 1:
        # every rank draws the same global batches (same seed, same sampler state) and keeps its shard
        train_loader = ShardedBatches(train_loader, rank, world)

    start_epoch = 0 if args.save_folder != args.load_folder else resume_from_epoch
    curr_prune_step = begin_prune_step = start_epoch * len(train_loader)
    end_prune_step = curr_prune_step + args.pruning_interval * len(train_loader)

    manager = Manager(args, model, shared_layer_info, masks, train_loader, val_loader, begin_prune_step, end_prune_step)
    if args.sync_free:
        from cpg_b200.train_loop import install_sync_free_train
        install_sync_free_train(manager)
    if args.mode == 'inference':
        manager.load_checkpoint_only_for_evaluate(resume_from_epoch, resume_folder)
        manager.validate(resume_from_epoch - 1)
        return _finish(world)

    # optimizer routing, :326-346
    sgd_params, adam_params = [], []
    head = '.{}.'.format(model.module.datasets.index(args.dataset))
    for name, param in model.named_parameters():
        if 'classifiers' in name:
            if head in name:
                sgd_params.append(param)
        elif 'piggymask' in name:
            adam_params.append(param)
        else:
            sgd_params.append(param)
    optimizers = Optimizers()
    # the same two optimizers (:339-346), each step one launch of cpg_b200.optim (bit-identical to torch's update;
    # CPGB_OPTIM=torch keeps the stock classes, as does a model that is not on the GPU)
    if os.environ.get('CPGB_OPTIM', 'cpgb') == 'cpgb' and sgd_params and sgd_params[0].is_cuda:
        from ..optim import SGD, Adam
    else:
        SGD, Adam = torch.optim.SGD, torch.optim.Adam
    optimizers.add(SGD(sgd_params, lr=args.lr, weight_decay=0.0, momentum=0.9, nesterov=True), args.lr)
    if adam_params:
        optimizers.add(Adam(adam_params, lr=args.lr_mask), args.lr_mask)
        if hasattr(optimizers.optimizers[-1], 'emit_packed_masks'):
            optimizers.optimizers[-1].emit_packed_masks(model)
    manager.load_checkpoint(optimizers, resume_from_epoch, resume_folder)
    if world > 1:
        # the all-reduce rides on the pruner call Manager.train makes right after backward()
        manager.pruner.grad_reducer = ddp.GradAllReducer(model, world)
    curr_lrs = [opt.param_groups[0]['lr'] for opt in optimizers]

    def record():
        path = args.pruning_ratio_to_acc_record_file
        if path and os.path.isfile(path):
            with open(path, 'r') as fh:
                return json.load(fh)
        return {}

    def must_prune_ratio():
        """:366-377 / :489-501: at the width cap with the goal still missed, the task may keep only its share.
        None when that situation does not apply (the reference only compares inside the branch: for the last task
        the share is the whole remainder, the ratio 0.0, and every sparsity "reaches" it)."""
        if (args.network_width_multiplier == args.max_allowed_network_width_multiplier
                and record()['0.0'] < baseline_acc):
            logging.info('we reach the upperbound and still do not get the accuracy over our target on curr task')
            remain = args.total_num_tasks - len(dataset_history)
            allow = round(1.0 / (remain + 1), 1)
            logging.info('remain_num_tasks: {}'.format(remain))
            logging.info('ratio_allow_for_curr_task: {:.4f}'.format(allow))
            return 1.0 - allow
        return None

    if args.mode == 'prune':
        if 'gradual_prune' in args.load_folder and args.save_folder == args.load_folder:
            args.epochs = 20 + resume_from_epoch
        logging.info('')
        logging.info('Before pruning: ')
        logging.info('Sparsity range: {} -> {}'.format(args.initial_sparsity, args.target_sparsity))
        limit = must_prune_ratio()
        if limit is not None and args.initial_sparsity >= limit:
            _finish(world, 6)
        manager.validate(start_epoch - 1)
        logging.info('')
    elif args.mode == 'finetune':
        if not args.finetune_again:
            manager.pruner.make_finetuning_mask()
            logging.info('Finetune stage...')
        else:
            logging.info('Piggymask Retrain...')
            best_retrain_acc = manager.validate(start_epoch - 1)
            stale_epochs = 0
        stop_lr_mask = True
        if manager.pruner.calculate_curr_task_ratio() == 0.0:
            logging.info('There is no left space in convolutional layer for curr task'
                         ', we will try to use prior experience as long as possible')
            stop_lr_mask = False

    avg_train_acc = avg_val_acc = 0.0
    epoch_idx = start_epoch - 1
    for epoch_idx in range(start_epoch, args.epochs):
        avg_train_acc, curr_prune_step = manager.train(optimizers, epoch_idx, curr_lrs, curr_prune_step)
        avg_train_acc = _all_mean(avg_train_acc, world, device)
        avg_val_acc = manager.validate(epoch_idx)
        if args.finetune_again:
            if avg_val_acc > best_retrain_acc:
                best_retrain_acc, stale_epochs = avg_val_acc, 0
                if rank == 0:
                    for path in os.listdir(args.save_folder):
                        if '.pth.tar' in path:
                            os.remove(os.path.join(args.save_folder, path))
                    manager.save_checkpoint(optimizers, epoch_idx, args.save_folder)
            else:
                stale_epochs += 1
            if stale_epochs == 5:
                logging.info('stop retraining')
                _finish(world, 0)
        if args.mode == 'finetune':
            if epoch_idx + 1 in (50, 80):
                for group in optimizers[0].param_groups:
                    group['lr'] *= 0.1
                curr_lrs[0] = optimizers[0].param_groups[0]['lr']
            if len(optimizers.lrs) == 2:
                if epoch_idx + 1 == 50:
                    for group in optimizers[1].param_groups:
                        group['lr'] *= 0.2
                if stop_lr_mask and epoch_idx + 1 == 70:
                    for group in optimizers[1].param_groups:
                        group['lr'] *= 0.0
                curr_lrs[1] = optimizers[1].param_groups[0]['lr']

    if avg_train_acc > 0.95 and rank == 0:
        manager.save_checkpoint(optimizers, epoch_idx, args.save_folder)
    logging.info('-' * 16)

    code = None
    if args.pruning_ratio_to_acc_record_file:
        data = record()
        if args.mode == 'finetune' and not args.test_piggymask:
            data[0.0] = round(avg_val_acc, 4)
            if rank == 0:
                with open(args.pruning_ratio_to_acc_record_file, 'w') as fh:
                    json.dump(data, fh)
            no_space = manager.pruner.calculate_curr_task_ratio() == 0.0
            if avg_train_acc > 0.95 and avg_val_acc >= baseline_acc:
                if no_space:
                    logging.info('There is no left space in convolutional layer for curr task, so needless to prune')
                    code = 5
            elif (args.network_width_multiplier == args.max_allowed_network_width_multiplier
                  and avg_val_acc < baseline_acc):
                code = 5 if no_space else 0
            else:
                logging.info("It's time to expand the Network")
                logging.info('Auto expand network')
                code = 2
        elif args.mode == 'prune':
            if avg_train_acc > 0.95:
                data[args.target_sparsity] = round(avg_val_acc, 4)
                if rank == 0:
                    with open(args.pruning_ratio_to_acc_record_file, 'w') as fh:
                        json.dump(data, fh)
                if world > 1:
                    dist.barrier()
                limit = must_prune_ratio()
                if limit is not None and args.target_sparsity >= limit:
                    code = 6
            else:
                code = 6
    return _finish(world, code)


def _finish(world, code=None):
    """Leave the process group cleanly, then exit with the reference's code (None: fall off the end, exit 0)."""
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    if code is not None:
        sys.exit(code)
    return 0


if __name__ == '__main__':
    main()
