"""Drop-in for the reference's ``transport`` package (LDMAE/transport/__init__.py:3-72)."""
from .transport import ModelType, PathType, Sampler, Transport, WeightType, create_transport  # noqa: F401
