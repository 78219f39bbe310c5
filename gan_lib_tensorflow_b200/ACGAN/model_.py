"""ACGAN/model_.py: the narrow generator variant of ACGAN/model.py -- a 4x4x128 seed and 128-channel 'up' blocks
(model_.py:33-42); the discriminator and everything else are identical."""
from __future__ import annotations

from .model import ACGAN as _ACGAN
from .model import discriminator_losses, generator_losses  # noqa: F401


class ACGAN(_ACGAN):
    G_INPUT_DIM = 128
    G_DIM = 128
