"""Counterpart of the reference's lib/models/util.py (util.py:2-25)."""


class AverageMeter(object):
    """Running average of a scalar (util.py:2-17): ``val`` last value, ``avg`` weighted mean."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.sum, self.count = 0, 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def set_requires_grad(nets, requires_grad):
    """Freeze / unfreeze every parameter of every net in ``nets`` (util.py:19-22).

    The CUDA backward honours it: a discriminator whose parameters are all frozen only runs the data-gradient
    kernels (no weight-gradient GEMMs), exactly the work autograd skips in the reference's G step
    (wgan_l1.py:169)."""
    for net in nets:
        for param in net.parameters():
            param.requires_grad = requires_grad


def count_parameters(model):
    """util.py:24-25."""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
