"""CPU oracle for the GAN-inpainting train step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the CPU baseline being timed), never as the thing shipped.

What it is: a plain fp32 PyTorch *functional* restatement (``torch.nn.functional``
over a flat ``{state_dict key: tensor}`` dict) of the reference's hot path:

* ``oracle.nets``    — UnetGenerator / PatchGANDiscriminator forward
                        (reference ``lib/models/networks.py:216-363``)
* ``oracle.losses``  — RMSELoss / LocalLoss (reference ``lib/models/loss.py:11-47``)
* ``oracle.steps``   — the per-batch train-step recipes
                        (reference ``experiment_list/wgan_l1.py:110-186``,
                        ``minimaxgan_rmse.py:110-173``,
                        ``experiment1_global_local_D.py:139-200``)
* ``oracle.ssim``    — the SSIM metric (reference ``lib/pytorch_ssim/__init__.py:10-40``)
* ``oracle.synth``   — the deterministic synthetic inputs of SURVEY.md §8(d)

Parity pinning: the reference ships **no** tests or golden vectors (SURVEY.md
D6), and all arithmetic below the ``nn.Module`` line is third-party PyTorch
(unpinned by the reference; torch 2.11.0 here).  The oracle is therefore pinned
against *outputs of the reference itself run in the build container*:
``oracle/pin_against_reference.py`` imports ``/root/reference/lib/models``
(read-only mount), loads oracle-initialised weights into the reference's own
modules, checks the restatement agrees, and writes the small golden fixtures
in ``tests/golden/`` which travel to the GPU box (the reference mount does not).
``oracle/pin_ssim_against_reference.py`` and ``oracle/pin_tv_against_reference.py`` do the same for the SSIM metric and the
total-variation loss (``reference_ssim.pt``, ``reference_tv.pt``); ``oracle.steps.calculate_metric`` restates the evaluation pass
of ``lib/models/evaluate.py:87-177`` (that file does not parse, so it is pinned only through the loss classes it calls).
"""
