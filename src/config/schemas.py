"""Drop-in for the reference's ``src/config/schemas.py`` (checkpoints pickle these classes by module path)."""
from adafortitran_b200.config import BaseConfig, ModelConfig, OFDMParams, PilotParams, SystemConfig

__all__ = ["OFDMParams", "PilotParams", "SystemConfig", "BaseConfig", "ModelConfig"]
