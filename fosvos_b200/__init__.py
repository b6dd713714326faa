"""fosvos_b200: the OSVOS VGG-16 per-frame hot path of Fast-OSVOS, B200-native.

Public surface = the reference's own (``networks/osvos_vgg.py``, ``layers/osvos_layers.py``,
the VGG optimizer policy of ``util/network_provider.py`` and the loops of ``train_online.py`` /
``util/experiment_helper.py``), backed by hand-written sm_100a kernels behind a C ABI
(``include/fosvos_b200.h``).  No CPU fallback: ops raise when the library or a B200 is missing.
"""
from .layers import (center_crop, class_balanced_cross_entropy_loss, interp_surgery, l1_loss, logit, mse_loss, sigmoid_np,
                     upsample_filt)
from .networks import OSVOS_VGG
from .optim import FusedAdam, FusedSGD, get_optimizer_offline, get_optimizer_online
from .online import finetune, finetune_samples, infer_sequence, region_iou, sequences_for_rank

__all__ = ["OSVOS_VGG", "class_balanced_cross_entropy_loss", "center_crop", "upsample_filt", "interp_surgery",
           "logit", "sigmoid_np", "mse_loss", "l1_loss", "FusedSGD", "FusedAdam", "get_optimizer_online", "get_optimizer_offline", "finetune", "finetune_samples",
           "infer_sequence", "region_iou", "sequences_for_rank"]
