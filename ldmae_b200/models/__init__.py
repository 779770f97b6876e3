from .lightningdit import LightningDiT, LightningDiT_models  # noqa: F401
