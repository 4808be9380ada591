"""``begin(state, loaders)`` of the reference's experiment_list/wgan_l1.py:15-277 on the B200 train step."""
from . import _runner


def begin(state, loaders):
    return _runner.run(state, loaders, title="wgan_l1", recipe="wgan", recon="l1", schedule=_runner.wgan_g_schedule(update_g_every=5))
