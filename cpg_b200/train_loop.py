"""Sync-free twin of ``Manager.train`` (reference: utils/manager.py:39-100).

The reference's epoch loop reads three device scalars back per batch -- ``classification_accuracy(...).cpu()``
(utils/__init__.py:48-51), ``train_loss.avg.item()`` / ``train_accuracy.avg.item()`` for the progress bar, and a
``calculate_sparsity()`` that reduces every task mask (utils/manager.py:60, 77-88).  At a 1.5 ms training step each of
them is a full pipeline drain.  ``train_sync_free`` runs the SAME sequence of operations on the model, the pruner and
the optimizers (zero_grad -> forward -> loss -> backward -> do_weight_decay_and_make_grads_zero -> step ->
gradually_prune) and keeps the two running metrics on the device, accumulated with the reference's own fp32
arithmetic (``sum += val * num``, ``n += num``, ``avg = sum / n``): the progress bar is refreshed every
``postfix_every`` batches instead of every batch, and the values returned at the end of the epoch -- the average
training accuracy the caller logs, and the prune step counter -- are bit-identical to ``Manager.train``'s
(tests/test_train_loop_reference_cpu.py runs both against each other with the unmodified reference).

The only host reads left inside an epoch are the progress-bar refreshes and the one status read-back of a prune event
(the ``sys.exit(2)`` condition of utils/prune.py:38-42).

``validate_sync_free`` does the same for ``Manager.validate`` (utils/manager.py:103-152), whose progress bar costs two
metric read-backs and up to four mask reductions per batch.

Usage: ``install_sync_free_train(manager)`` right after ``Manager(...)`` is constructed; the rest of the reference's main
script (``manager.train(optimizers, epoch_idx, curr_lrs, curr_prune_step)``, ``manager.validate(epoch_idx)``) stays as it
is.
"""
import logging
import types

import torch

POSTFIX_EVERY = 50


def _accuracy_on_device(output, target):
    """classification_accuracy (utils/__init__.py:48-51) without its .cpu(): the mean of exact 0/1 values is the
    correctly rounded quotient of two integers on either device."""
    pred = output.max(1, keepdim=True)[1]
    return pred.eq(target.view_as(pred)).float().mean()


def train_sync_free(self, optimizers, epoch_idx, curr_lrs, curr_prune_step, postfix_every=None):
    """Drop-in for ``Manager.train`` (bound to a Manager instance by ``install_sync_free_train``)."""
    from tqdm import tqdm
    every = int(postfix_every if postfix_every is not None else getattr(self, 'postfix_every', POSTFIX_EVERY))
    self.model.train()

    loss_sum = acc_sum = count = None       # fp32 scalars on the model's device, the reference's Metric arithmetic
    total = len(self.train_loader)

    def averages():
        if count is None:
            nan = float('nan')
            return nan, nan
        return (loss_sum / count).item(), (acc_sum / count).item()

    with tqdm(total=total, desc='Train Ep. #{}: '.format(epoch_idx + 1), disable=False, ascii=True) as t:
        for batch_idx, (data, target) in enumerate(self.train_loader):
            if self.args.cuda:
                data, target = data.cuda(), target.cuda()

            optimizers.zero_grad()
            output = self.model(data)

            num = data.size(0)
            if count is None:
                loss_sum = torch.zeros((), dtype=torch.float32, device=output.device)
                acc_sum = torch.zeros((), dtype=torch.float32, device=output.device)
                count = torch.zeros((), dtype=torch.float32, device=output.device)
            if self.args.dataset != 'face_verification':
                acc_sum += _accuracy_on_device(output.detach(), target) * num

            loss = self.criterion(output, target)
            loss_sum += loss.detach() * num
            count += num
            loss.backward()

            # Set fixed param grads to 0.
            self.pruner.do_weight_decay_and_make_grads_zero()
            optimizers.step()

            if self.args.mode == 'prune':
                self.pruner.gradually_prune(curr_prune_step)
                curr_prune_step += 1

            if every > 0 and ((batch_idx + 1) % every == 0 or batch_idx + 1 == total):
                avg_loss, avg_acc = averages()
                t.set_postfix({'loss': avg_loss,
                               'accuracy': '{:.2f}'.format(100. * avg_acc),
                               'lr': curr_lrs[0],
                               'sparsity': self.pruner.calculate_sparsity(),
                               'network_width_mpl': self.args.network_width_multiplier})
            t.update(1)

    avg_loss, avg_acc = averages()
    if self.args.dataset == 'face_verification' and count is not None:
        avg_acc = float('nan')              # the reference never updates the metric there: 0 / 0
    summary = {'loss': '{:.3f}'.format(avg_loss),
               'accuracy': '{:.2f}'.format(100. * avg_acc),
               'lr': curr_lrs[0],
               'sparsity': '{:.3f}'.format(self.pruner.calculate_sparsity()),
               'network_width_mpl': self.args.network_width_multiplier}

    if self.args.log_path:
        logging.info(('In train()-> Train Ep. #{} '.format(epoch_idx + 1)
                      + ', '.join(['{}: {}'.format(k, v) for k, v in summary.items()])))
    return avg_acc, curr_prune_step


def validate_sync_free(self, epoch_idx, biases=None, postfix_every=None):
    """Drop-in for ``Manager.validate`` (utils/manager.py:103-152): the reference refreshes the progress bar after
    every batch with two metric read-backs and up to four mask statistics (``calculate_shared_part_ratio`` reduces every
    piggymask each time); here the metrics stay on the device and the bar is refreshed every ``postfix_every`` batches.
    Same ``apply_mask()`` first, same forward passes, same returned accuracy."""
    from tqdm import tqdm
    every = int(postfix_every if postfix_every is not None else getattr(self, 'postfix_every', POSTFIX_EVERY))
    self.pruner.apply_mask()
    self.model.eval()
    loss_sum = acc_sum = count = None
    total = len(self.val_loader)
    idx = self.inference_dataset_idx

    def averages():
        if count is None:
            nan = float('nan')
            return nan, nan
        return (loss_sum / count).item(), (acc_sum / count).item()

    with tqdm(total=total, desc='Val Ep. #{}: '.format(epoch_idx + 1), ascii=True) as t:
        with torch.no_grad():
            for batch_idx, (data, target) in enumerate(self.val_loader):
                if self.args.cuda:
                    data, target = data.cuda(), target.cuda()

                output = self.model(data)
                num = data.size(0)
                if count is None:
                    loss_sum = torch.zeros((), dtype=torch.float32, device=output.device)
                    acc_sum = torch.zeros((), dtype=torch.float32, device=output.device)
                    count = torch.zeros((), dtype=torch.float32, device=output.device)
                loss_sum += self.criterion(output, target) * num
                acc_sum += _accuracy_on_device(output, target) * num
                count += num

                if every > 0 and ((batch_idx + 1) % every == 0 or batch_idx + 1 == total):
                    avg_loss, avg_acc = averages()
                    postfix = {'loss': avg_loss,
                               'accuracy': '{:.2f}'.format(100. * avg_acc),
                               'sparsity': self.pruner.calculate_sparsity(),
                               'task{} ratio'.format(idx): self.pruner.calculate_curr_task_ratio()}
                    if idx != 1:
                        postfix['shared_ratio'] = self.pruner.calculate_shared_part_ratio()
                    postfix['zero ratio'] = self.pruner.calculate_zero_ratio()
                    postfix['mpl'] = self.args.network_width_multiplier
                    t.set_postfix(postfix)
                t.update(1)

    avg_loss, avg_acc = averages()
    summary = {'loss': '{:.3f}'.format(avg_loss),
               'accuracy': '{:.2f}'.format(100. * avg_acc),
               'sparsity': '{:.3f}'.format(self.pruner.calculate_sparsity()),
               'task{} ratio'.format(idx): '{:.3f}'.format(self.pruner.calculate_curr_task_ratio()),
               'zero ratio': '{:.3f}'.format(self.pruner.calculate_zero_ratio()),
               'mpl': self.args.network_width_multiplier}
    if idx != 1:
        summary['shared_ratio'] = '{:.3f}'.format(self.pruner.calculate_shared_part_ratio())

    if self.args.log_path:
        logging.info(('In validate()-> Val Ep. #{} '.format(epoch_idx + 1)
                      + ', '.join(['{}: {}'.format(k, v) for k, v in summary.items()])))
    return avg_acc


def install_sync_free_train(manager, postfix_every=POSTFIX_EVERY, validate=True):
    """Replace ``manager.train`` (and, with `validate`, ``manager.validate``) of an unmodified
    ``utils.manager.Manager`` by the sync-free loops."""
    manager.postfix_every = int(postfix_every)
    manager.train = types.MethodType(train_sync_free, manager)
    if validate:
        manager.validate = types.MethodType(validate_sync_free, manager)
    return manager
