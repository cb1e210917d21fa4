"""PyTorch host models that CALL the correlation hot path (they are not part of it).

The reference keeps its CCE encoder, GRU update block and losses in PyTorch; so does this
repo.  `FocusRAFT` keeps the reference's state_dict key names so `ffraft_*.pth` checkpoints
load unchanged.
"""
from .focusraft import FocusRAFT, RAFTBody, build_focusraft  # noqa: F401
from .focuspwc import FocusPWC, backwarp  # noqa: F401
from .losses import CPCL, EPELoss, MixLoss, build_losses  # noqa: F401
from .train import TrainStep, parallel_model  # noqa: F401
