"""Host-side training/checkpoint helpers the reference's tools import from `tree_learn.util`
(reference tree_learn/util/train.py): same names and behaviour, written for this package.  Plain torch -- none of
this is on the GPU hot path."""
import functools
import os

import torch
from torch.utils.data import DataLoader

from .model import point_wise_loss  # noqa: F401  (re-exported under the reference's name)


def is_multiple(num, multiple):
    return num != 0 and num % multiple == 0


def cuda_cast(func):
    """Decorator: move every tensor argument to the current CUDA device (reference util/train.py:28-43)."""
    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        move = lambda v: v.cuda() if isinstance(v, torch.Tensor) else v   # noqa: E731
        return func(*[move(a) for a in args], **{k: move(v) for k, v in kwargs.items()})
    return wrapper


def checkpoint_save(epoch, model, optimizer, work_dir, save_freq=1):
    """Write epoch_{n}.pth = {net (cpu), optimizer, epoch}; drop epoch_{n-1}.pth unless n-1 is a multiple of save_freq."""
    net = model.module if hasattr(model, 'module') else model
    state = {'net': {k: v.cpu() for k, v in net.state_dict().items()}, 'optimizer': optimizer.state_dict(), 'epoch': epoch}
    torch.save(state, os.path.join(work_dir, f'epoch_{epoch}.pth'))
    prev = os.path.join(work_dir, f'epoch_{epoch - 1}.pth')
    if os.path.isfile(prev) and not is_multiple(epoch - 1, save_freq):
        os.remove(prev)


def load_checkpoint(checkpoint, logger, model, optimizer=None, strict=False):
    """Non-strict load of state_dict['net']; keys whose shapes differ are dropped (reference util/train.py:65-102),
    which is how the SoftGroup/HAIS pre-training checkpoint with another input width is accepted.  Returns epoch+1."""
    net = model.module if hasattr(model, 'module') else model
    state = torch.load(checkpoint, map_location='cpu')
    src = state['net'] if 'net' in state else state
    own = net.state_dict()
    keep = {k: v for k, v in src.items() if k not in own or tuple(own[k].shape) == tuple(v.shape)}
    dropped = sorted(set(src) - set(keep))
    missing, unexpected = net.load_state_dict(keep, strict=strict)
    if logger is not None:
        if dropped:
            logger.info(f'removed keys in source state_dict due to size mismatch: {", ".join(dropped)}')
        if missing:
            logger.info(f'missing keys in source state_dict: {", ".join(missing)}')
        if unexpected:
            logger.info(f'unexpected key in source state_dict: {", ".join(unexpected)}')
    if optimizer is not None and 'optimizer' in state:
        optimizer.load_state_dict(state['optimizer'])
    return state.get('epoch', 0) + 1


def build_optimizer(model, optim_cfg):
    cfg = dict(optim_cfg)
    kind = cfg.pop('type')
    return getattr(torch.optim, kind)(filter(lambda p: p.requires_grad, model.parameters()), **cfg)


def build_dataloader(dataset, batch_size=1, num_workers=1, training=True, dist=False):
    """Reference util/train.py:125-141: a DistributedSampler only for `dist and training`; shuffle / drop_last in training."""
    sampler = torch.utils.data.distributed.DistributedSampler(dataset, shuffle=True) if (dist and training) else None
    return DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, collate_fn=dataset.collate_fn,
                      shuffle=(training and sampler is None), sampler=sampler, drop_last=training, pin_memory=True)
