"""Drop-in for the reference's ``src/models/__init__.py:1-3``."""
from adafortitran_b200.estimators import AdaFortiTranEstimator, BaseFortiTranEstimator, FortiTranEstimator, LinearEstimator

__all__ = ["FortiTranEstimator", "AdaFortiTranEstimator", "LinearEstimator", "BaseFortiTranEstimator"]
