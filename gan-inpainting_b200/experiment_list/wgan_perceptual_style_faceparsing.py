"""``begin(state, loaders)`` of the reference's experiment_list/wgan_perceptual_style_faceparsing.py:15-349 on the B200 train step.

The reference script does not run (``unique_labels`` :117 and ``recon_loss`` :221 are undefined, ``tv_loss`` :219 is unqualified,
``LocalLoss(loss.RMSELoss())`` :63 cannot be constructed).  This is the step it describes: the ``wgan_l1`` schedule
with ``g_loss`` = adversarial + RMSE + local reconstruction (the usable ``LocalLoss`` base closest to its intent: masked,
micro-averaged MSE) + ``tv_loss(inpainted, 1)`` + ``perceptual_and_style_loss(inpainted, ground, 0.01, 0.01)`` (:208-221).

NOT computed: the face-parsing term (:212-213) — it needs ``state['segmentation_model']``, a separately trained ngf = 32 U-Net whose
checkpoint does not exist here.  A ``segmentation_model`` in ``state`` raises unless
``state['ignore_segmentation_model']`` is set, so that nobody trains without that term unknowingly.  The VGG19 of the feature
losses is ``lib.models.loss.vgg`` (``loss.set_vgg``; random initialisation unless trained weights are installed).
"""
from . import _runner

LOSS_TERMS = dict(tv_weight=1.0, feature_weights=(0.01, 0.01), local_recon="mse")  # :209, :216, :219


def begin(state, loaders):
    if state.get("segmentation_model") is not None and not state.get("ignore_segmentation_model", False):
        raise NotImplementedError("the face-parsing loss term of wgan_perceptual_style_faceparsing.py:212-213 is not built; "
                                  "set state['ignore_segmentation_model'] = True to train without it")
    return _runner.run(state, loaders, title="wgan_perceptual_style_faceparsing", recipe="wgan", recon="rmse",
                       schedule=_runner.wgan_g_schedule(update_g_every=5), loss_terms=LOSS_TERMS)
