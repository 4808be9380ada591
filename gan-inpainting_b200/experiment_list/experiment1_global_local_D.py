"""The global + local discriminator recipe of the reference's experiment_list/experiment1_global_local_D.py:139-200
(a standalone argparse script there; exposed here with the same ``begin(state, loaders)`` surface as the other recipes).
``state`` may carry ``lambda_global`` / ``lambda_local`` (defaults 300 / 300, experiment1_global_local_D.py:26-27)."""
from . import _runner


def begin(state, loaders):
    lambdas = (float(state.get("lambda_global", 300.0)), float(state.get("lambda_local", 300.0)))
    return _runner.run(state, loaders, title="experiment1_global_local_D", recipe="global_local", recon="rmse", schedule=_runner.every_batch,
                       local_d=True, lambdas=lambdas)
