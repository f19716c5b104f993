"""Drop-in for the reference's ``src/models/adafortitran.py`` (:5-22)."""
from adafortitran_b200.estimators import AdaFortiTranEstimator

__all__ = ["AdaFortiTranEstimator"]
