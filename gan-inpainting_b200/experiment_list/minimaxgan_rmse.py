"""``begin(state, loaders)`` of the reference's experiment_list/minimaxgan_rmse.py:15-260 on the B200 train step."""
from . import _runner


def begin(state, loaders):
    return _runner.run(state, loaders, title="minimaxgan_rmse", recipe="minimax", recon="rmse", schedule=_runner.every_batch)
