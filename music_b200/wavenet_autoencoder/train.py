"""Drop-in for the reference's `wavenet_autoencoder/train.py`: optimizer factory, checkpoint helpers, one train step.

Reference: get_optimizer :26-34, save_model :36-41, load_model :44-64, hot loop :117-134
(zero_grad -> net(music_piece) -> CrossEntropyLoss(outputs, target.view(-1)) -> backward -> step).  The loss is
applied to the module's softmax OUTPUT (a second softmax inside CrossEntropyLoss), exactly as in wavenet/train.py.
Forward and backward of the module run in libwavenet_b200.so (csrc/ae.cu); the optimizers are torch.optim objects
on the module's parameters, as in the reference.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.optim as optim

from .model1 import wavenet_autoencoder  # noqa: F401  (re-exported like the reference's `from model1 import ...`)


def get_optimizer(model, optimizer_type, learning_rate, momentum1=False):
    """Names as the reference spells them (:26-34).  Its 'sgd' branch calls the non-existent `optim.sgd`; SGD is meant."""
    if optimizer_type == 'sgd':
        return optim.SGD(model.parameters(), lr=learning_rate, momentum=momentum1)
    if optimizer_type == 'RMSprop':
        return optim.RMSprop(model.parameters(), lr=learning_rate, momentum=momentum1)
    if optimizer_type == 'Adam':
        return optim.Adam(model.parameters(), lr=learning_rate)
    if optimizer_type == 'lbfgs':
        return optim.LBFGS(model.parameters(), lr=learning_rate)


def save_model(model, num_epoch, path):
    model_name = 'wavenet_autoencoder' + str(num_epoch) + '.model'
    checkpoint_path = path + model_name
    print('Storing checkpoint to {}...'.format(path))
    torch.save(OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items()), checkpoint_path)
    print('done')


def load_model(model, path, model_name):
    """Returns the model, or None when the file is missing; strips DataParallel's `module.` prefix (:44-64)."""
    checkpoint_path = path + model_name
    print("Trying to restore saved checkpoint from ", "{}".format(checkpoint_path))
    if not os.path.exists(checkpoint_path):
        print("No checkpoint found!")
        return None
    print("Checkpoint found, restoring!")
    state_dict = torch.load(checkpoint_path, map_location="cpu")
    keys = list(state_dict.keys())
    if keys[0][:6] == 'module':
        state_dict = OrderedDict((k[7:], v) for k, v in state_dict.items())
    model.load_state_dict(state_dict)
    return model


def all_reduce_grads_(params, dist=None, group=None):
    """Data-parallel gradient exchange for one process per GPU (the reference wraps the net in nn.DataParallel,
    train.py:93-97): ONE all-reduce of the flattened gradients, averaged over ranks.  Every rank holds the same number of
    loss rows, so this equals the gradient of the mean loss over the gathered batch.  Works on any backend (NCCL on the
    GPUs, gloo in the CPU tests)."""
    if dist is None:
        import torch.distributed as dist
    params = [p for p in params if p.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def train_step(net, optimizer, music_piece, target_piece, loss_func=None, cond_weights=None, distributed=False):
    """One iteration of the reference loop (:117-134).  `music_piece` (B,Q,L) float on the GPU, `target_piece` any
    integer tensor with B*W entries.  Returns the loss as a 0-d device tensor (no host sync; the reference's
    `loss.data[0]` is the caller's choice).  distributed=True averages the gradients over the process group first."""
    loss_func = loss_func or nn.CrossEntropyLoss()
    optimizer.zero_grad()
    outputs = net(music_piece, cond_weights=cond_weights)
    loss = loss_func(outputs, target_piece.reshape(-1).to(outputs.device, torch.int64))
    loss.backward()
    if distributed:
        all_reduce_grads_(net.parameters())
    optimizer.step()
    return loss.detach()
