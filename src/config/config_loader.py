"""Drop-in for the reference's ``src/config/config_loader.py`` (:16-78)."""
from adafortitran_b200.config import ConfigLoader, load_config

__all__ = ["ConfigLoader", "load_config"]
