"""``begin(state, loaders)`` of the reference's experiment_list/wgan_rmse.py (same loop, RMSE reconstruction; the reference file has a SyntaxError at :104) on the B200 train step."""
from . import _runner


def begin(state, loaders):
    return _runner.run(state, loaders, title="wgan_rmse", recipe="wgan", recon="rmse", schedule=_runner.wgan_g_schedule(update_g_every=5))
