"""Drop-in for the reference's ``src/config/__init__.py``."""
from adafortitran_b200.config import load_config

__all__ = ["load_config"]
