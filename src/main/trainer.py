"""Drop-in for the EVALUATION part of the reference's ``src/main/trainer.py``: ``ModelEvaluator`` (:259-347) and
``MODEL_REGISTRY`` (:381-385) over the B200 estimators.  ``ModelTrainer`` / ``TrainingLoop`` / ``train`` are not provided."""
from adafortitran_b200.evaluate import MODEL_REGISTRY, ModelEvaluator, load_checkpoint

__all__ = ["ModelEvaluator", "MODEL_REGISTRY", "load_checkpoint"]
