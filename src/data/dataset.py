"""Drop-in for the reference's ``src/data/dataset.py``."""
from adafortitran_b200.data import MatDataset, extract_pilots
from adafortitran_b200.evaluate import get_test_dataloaders

__all__ = ["MatDataset", "get_test_dataloaders", "extract_pilots"]
