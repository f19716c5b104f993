"""Drop-in for the helpers of the reference's ``src/utils.py`` that the EVALUATION path uses (src/utils.py:68-110,164-180,
233-265).  The plotting / early-stopping / model-table helpers of that file belong to the training loop, which is out of
scope here: importing them from this shim raises AttributeError rather than pretending."""
import numpy as np
import torch

from adafortitran_b200.data import extract_values

__all__ = ["extract_values", "concat_complex_channel", "to_db", "mse"]


def concat_complex_channel(channel_matrix: torch.Tensor) -> torch.Tensor:
    """Real view of a complex channel matrix: real and imaginary parts concatenated along dim 1 (src/utils.py:164-180)."""
    return torch.cat((torch.real(channel_matrix), torch.imag(channel_matrix)), dim=1)


def to_db(val):
    """10 log10(val) (src/utils.py:233-245)."""
    return 10 * np.log10(val)


def mse(x, y):
    """MSE in dB between two complex numpy arrays (src/utils.py:248-265)."""
    return to_db(np.mean(np.square(np.abs(x - y))))
