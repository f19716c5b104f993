"""Drop-in for the reference's ``src/data`` (src/data/dataset.py:36-262): the .mat dataset and the test-set loaders."""
from adafortitran_b200.data import MatDataset
from adafortitran_b200.evaluate import get_test_dataloaders

__all__ = ["MatDataset", "get_test_dataloaders"]
