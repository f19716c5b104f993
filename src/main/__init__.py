"""Evaluation half of the reference's ``src/main`` package (src/main/trainer.py:259-347): ``ModelEvaluator`` and the
model registry.  The training loop (``ModelTrainer``, ``TrainingLoop``, callbacks, ``train``) is out of scope."""
