"""adafortitran_b200 -- B200-native (sm_100a) inference forward pass of AdaFortiTran / FortiTran.

Public surface (mirrors the reference's ``src.models`` / ``src.config``):
    FortiTranEstimator, AdaFortiTranEstimator, BaseFortiTranEstimator
    SystemConfig, ModelConfig, load_config
"""
from .config import ConfigLoader, ModelConfig, OFDMParams, PilotParams, SystemConfig, load_config
from .estimators import AdaFortiTranEstimator, BaseFortiTranEstimator, FortiTranEstimator

__all__ = [
    "AdaFortiTranEstimator", "BaseFortiTranEstimator", "FortiTranEstimator",
    "SystemConfig", "ModelConfig", "OFDMParams", "PilotParams", "ConfigLoader", "load_config",
]
__version__ = "0.1.0"
