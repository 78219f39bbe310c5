"""`lib.ops.pixelnorm.Pixelnorm` is called by PGGAN/model_nvidia.py:63,68 but does not exist in the reference;
it is provided here as the alias of normalization.pixel_norm."""
from .normalization import pixel_norm as Pixelnorm  # noqa: F401
