"""Neural Jump ODE -- B200-native (sm_100a) drop-in for the hot path of
alexander-dybdahl/neural-jump-ode: same import surface as the reference package
(reference: neural_jump_ode/__init__.py:3-6)."""

from .models.jump_ode import NeuralJumpODE, nj_ode_loss
from .packed import PackedBatch
from .optim import FlatAdam
from .training import train_epoch_packed, validate_packed, relative_loss_packed

__version__ = "0.1.0"
__all__ = ["NeuralJumpODE", "nj_ode_loss", "PackedBatch", "FlatAdam", "train_epoch_packed", "validate_packed", "relative_loss_packed"]
