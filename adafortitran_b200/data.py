"""Input side of the hot path: what ``MatDataset`` / ``extract_values`` hand to the model
(reference src/data/dataset.py:36-190, src/utils.py:68-110), with the per-file pilot extraction batched on the GPU.

* :func:`extract_values` -- file name -> the 6-tuple of metadata the reference collates into ``meta_data``.
* :func:`extract_pilots` -- batched ``_process_channel_data``: non-zero entries of the sparse LS grid, row-major
  (``aft_extract_pilots``; device tensors in, device tensors out, ``ValueError`` on a wrong pilot count).
* :class:`MatDataset` -- same ``__getitem__`` contract as the reference for ``.mat`` files holding ``H`` [scs, symbols, >=2];
  :func:`collate_on_device` stacks raw grids of a batch and extracts all pilots in one launch.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _capi
from .config import PilotParams

_NAME = re.compile(r"(\d+)_SNR-(\d+)_DS-(\d+)_DOP-(\d+)_N-(\d+)_([A-Z\-]+)\.mat")


def extract_values(file_name: str):
    """``'{n}_SNR-{snr}_DS-{ds}_DOP-{dop}_N-{pilot_freq}_{channel}.mat'`` -> five float32 tensors of shape [1] and
    ``[channel]`` (reference src/utils.py:68-110, ``re.match`` semantics); ``ValueError`` otherwise."""
    m = _NAME.match(file_name)
    if not m:
        raise ValueError("Cannot extract file information.")
    vals = tuple(torch.tensor([int(m.group(i))], dtype=torch.float) for i in range(1, 6))
    return (*vals, [m.group(6)])


def extract_pilots(ls_grid: torch.Tensor, pilot_size: Tuple[int, int]) -> torch.Tensor:
    """complex64 ``[B, scs, symbols]`` sparse LS grids on a CUDA device -> complex64 ``[B, pilot_scs, pilot_symbols]``.
    Every sample must hold exactly ``pilot_scs * pilot_symbols`` non-zero entries (dataset.py:129-133)."""
    if ls_grid.dim() != 3 or not ls_grid.is_complex():
        raise ValueError(f"expected a complex [B, subcarriers, symbols] tensor, got {tuple(ls_grid.shape)} {ls_grid.dtype}")
    if ls_grid.device.type != "cuda":
        raise RuntimeError("adafortitran_b200 has no CPU fallback: extract_pilots needs a CUDA tensor")
    grid = ls_grid.to(torch.complex64).contiguous()
    batch, cells, expected = grid.shape[0], grid.shape[1] * grid.shape[2], pilot_size[0] * pilot_size[1]
    pilots = torch.zeros((batch, pilot_size[0], pilot_size[1]), dtype=torch.complex64, device=grid.device)
    counts = torch.empty((batch,), dtype=torch.int32, device=grid.device)
    with torch.cuda.device(grid.device):
        stream = torch.cuda.current_stream(grid.device).cuda_stream
        _capi.check(_capi.lib().aft_extract_pilots(C.c_void_p(grid.data_ptr()), C.c_void_p(pilots.data_ptr()),
                                                   C.c_void_p(counts.data_ptr()), batch, cells, expected, C.c_void_p(stream)))
    bad = (counts != expected).nonzero()
    if bad.numel():
        i = int(bad[0])
        raise ValueError(f"Error processing channel data: Expected {expected} pilot values, got {int(counts[i])} (sample {i})")
    return pilots


class MatDataset(torch.utils.data.Dataset):
    """Reference ``MatDataset`` contract (dataset.py:36-190): item = (LS pilots [pilot_scs, pilot_symbols] c64, ground
    truth [scs, symbols] c64, metadata 6-tuple).  ``raw=True`` returns the sparse LS grid instead of the pilots so that
    :func:`collate_on_device` can extract a whole batch on the GPU."""

    def __init__(self, data_dir: Union[str, Path], pilot_params: Union[PilotParams, Sequence[int]],
                 transform: Optional[Callable] = None, raw: bool = False) -> None:
        self.data_dir = Path(data_dir)
        if not hasattr(pilot_params, "num_scs"):   # convenience: (pilot_scs, pilot_symbols)
            pilot_params = PilotParams(num_scs=pilot_params[0], num_symbols=pilot_params[1])
        self.pilot_params = pilot_params
        self.transform = transform
        self.raw = raw
        if not self.data_dir.exists():
            raise FileNotFoundError(f"Data directory not found: {self.data_dir}")
        self.file_list: List[Path] = list(self.data_dir.glob("*.mat"))
        if not self.file_list:
            raise ValueError(f"No .mat files found in {self.data_dir}")

    def __len__(self) -> int:
        return len(self.file_list)

    def _grids(self, path: Path):
        import scipy.io as sio
        mat = sio.loadmat(path)
        if "H" not in mat or mat["H"].shape[-1] < 2:
            raise ValueError("Invalid .mat file format: missing required data")
        return (torch.tensor(mat["H"][:, :, 0], dtype=torch.cfloat), torch.tensor(mat["H"][:, :, 1], dtype=torch.cfloat))

    def __getitem__(self, idx: int):
        if not 0 <= idx < len(self):
            raise IndexError(f"Index {idx} out of range for dataset of size {len(self)}")
        path = self.file_list[idx]
        try:
            h_ideal, ls = self._grids(path)
            if self.raw:
                h_est = ls
            else:   # host form of the extraction, one file at a time (dataset.py:118-139)
                h_est = ls[ls != 0]
                expected = self.pilot_params.num_scs * self.pilot_params.num_symbols
                if h_est.numel() != expected:
                    raise ValueError(f"Expected {expected} pilot values, got {h_est.numel()}")
                h_est = h_est.view(self.pilot_params.num_scs, self.pilot_params.num_symbols)
            meta = extract_values(path.name)
            if self.transform:
                h_est, h_ideal = self.transform(h_est), self.transform(h_ideal)
            return h_est, h_ideal, meta
        except Exception as e:
            raise ValueError(f"Error processing file {path}: {e}")


def collate_on_device(batch, pilot_size: Tuple[int, int], device: Union[str, torch.device] = "cuda"):
    """Collate items of a ``MatDataset(raw=True)``: one host->device copy of the stacked sparse grids, then one
    ``aft_extract_pilots`` launch.  Returns (pilots [B, ps, pt] c64 on `device`, truth [B, scs, symbols] c64 on
    `device`, metadata collated like torch's default collate: five float32 [B, 1] tensors and [tuple of channel names])."""
    grids = torch.stack([b[0] for b in batch]).to(device, non_blocking=True)
    truth = torch.stack([b[1] for b in batch]).to(device, non_blocking=True)
    meta = [torch.stack([b[2][k] for b in batch]) for k in range(5)] + [[tuple(b[2][5][0] for b in batch)]]
    return extract_pilots(grids, pilot_size), truth, tuple(meta)
