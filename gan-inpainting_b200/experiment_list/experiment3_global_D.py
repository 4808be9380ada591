"""The global-discriminator-only recipe of the reference's experiment_list/experiment3_global_D.py:140-196 (a standalone argparse
script there; exposed here with the same ``begin(state, loaders)`` surface as the other recipes): G step first on the un-composited
generator output (mask not ceil'ed), BCE heads on a sigmoid PatchGAN, ``g_loss = BCE(D(G(x)), 1) + lambda1 * RMSE``, then one D
backward of ``BCE(D(ground), 1) + BCE(D(G(x).detach()), 0)``; Adam(2e-4, (0.5, 0.999)) for both networks (:121-122); SSIM logged per
epoch (:201).  ``state`` may carry ``lambda_global`` (default 300, :26)."""
from . import _runner


def begin(state, loaders):
    lambdas = (float(state.get("lambda_global", 300.0)), 0.0)
    return _runner.run(state, loaders, title="experiment3_global_D", recipe="global_d", recon="rmse", schedule=_runner.every_batch, lambdas=lambdas,
                       log_ssim=True)
