"""``begin(state, loaders)`` of the reference's experiment_list/minimaxgan_l1.py:15-260 on the B200 train step."""
from . import _runner


def begin(state, loaders):
    return _runner.run(state, loaders, title="minimaxgan_l1", recipe="minimax", recon="l1", schedule=_runner.every_batch)
