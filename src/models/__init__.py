"""Drop-in for the reference's ``src/models/__init__.py:1-3`` (hot-path estimators only)."""
from adafortitran_b200.estimators import AdaFortiTranEstimator, BaseFortiTranEstimator, FortiTranEstimator

__all__ = ["FortiTranEstimator", "AdaFortiTranEstimator", "BaseFortiTranEstimator"]
