"""Evaluation loop around the hot path: the reference's ``ModelEvaluator`` /
``get_test_dataloaders`` / checkpoint format (src/main/trainer.py:259-347,648-703, src/data/dataset.py:193-262) with the
per-batch work on the device: pilots are extracted from the raw LS grids in one launch (``data.collate_on_device``) and
the squared-error sums stay on the GPU (``aft_error_sums``, fp64) -- one host read per data loader instead of one
``.item()`` per batch.

CLI:  python -m adafortitran_b200.evaluate --system_config_path config/system_config.yaml \\
          --model_config_path config/adafortitran.yaml --model_name adafortitran --test_set data/test/DS_test_set \\
          [--checkpoint ckpt.pt] [--precision bf16] [--batch_size 512]
prints one JSON object {"<value>": mse_db, ...} per test set (sub-directories named VAR_value, e.g. DS_50, SNR_10).
"""
from __future__ import annotations

import ctypes as C
import functools
import json
import logging
import math
from pathlib import Path
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch.utils.data import DataLoader

from . import _capi, data
from .config import PilotParams, load_config
from .estimators import AdaFortiTranEstimator, FortiTranEstimator, LinearEstimator

MODEL_REGISTRY = {"linear": LinearEstimator, "fortitran": FortiTranEstimator, "adafortitran": AdaFortiTranEstimator}


def to_db(x: float) -> float:
    """10 log10(x) (reference src/utils.py:233-245)."""
    return 10.0 * math.log10(x)


def get_test_dataloaders(dataset_dir: Union[str, Path], pilot_params: PilotParams, batch_size: int,
                         device: Union[str, torch.device] = "cuda") -> List[Tuple[str, DataLoader]]:
    """One loader per sub-directory of `dataset_dir` (reference dataset.py:193-262: no shuffling, no workers).  Batches
    are collated on `device`: (pilots [B, ps, pt] c64, truth [B, scs, symbols] c64, metadata) -- the same triple the
    reference's default collate yields, already on the GPU."""
    dataset_dir = Path(dataset_dir)
    if not dataset_dir.exists():
        raise FileNotFoundError(f"Dataset directory not found: {dataset_dir}")
    subdirs = [d for d in dataset_dir.iterdir() if d.is_dir()]
    if not subdirs:
        raise ValueError(f"No subdirectories found in {dataset_dir}")
    pilot_size = (pilot_params.num_scs, pilot_params.num_symbols)
    collate = functools.partial(data.collate_on_device, pilot_size=pilot_size, device=device)
    return [(d.name, DataLoader(data.MatDataset(d, pilot_params, raw=True), batch_size=batch_size, shuffle=False, num_workers=0,
                                collate_fn=collate)) for d in subdirs]


class ModelEvaluator:
    """Reference ``ModelEvaluator`` (trainer.py:259-347) over the B200 estimators."""

    def __init__(self, model, device: Union[str, torch.device], logger: Optional[logging.Logger] = None):
        self.model = model
        self.device = torch.device(device)
        self.logger = logger or logging.getLogger(__name__)

    def _forward_pass(self, coarse_estimated_channel: torch.Tensor, model, meta_data: Optional[Tuple] = None) -> torch.Tensor:
        if isinstance(model, AdaFortiTranEstimator):
            if meta_data is None:
                raise ValueError("AdaFortiTranEstimator requires meta_data but it was not provided")
            return model(coarse_estimated_channel, meta_data)
        return model(coarse_estimated_channel)          # Linear and FortiTran models don't use meta_data

    def _evaluate_dataloader(self, dataloader, loss_fn=None) -> float:
        """Mean squared error per complex grid element over the loader: exactly the reference's
        ``sum_b 2 * MSELoss(cat(re, im)) * B_b / sum_b B_b`` (trainer.py:328-347); `loss_fn` is accepted for signature
        compatibility and ignored (the reference always passes ``nn.MSELoss()``)."""
        self.model.eval()
        sums = torch.zeros(2, dtype=torch.float64, device=self.device)
        elements = 0
        with torch.no_grad(), torch.cuda.device(self.device):
            for estimated_channel_input, ideal_channel, meta_data in dataloader:
                est = self._forward_pass(estimated_channel_input, self.model, meta_data)
                if not est.is_complex():                       # LinearEstimator is real-valued
                    est = est.to(torch.complex64)
                est = est.to(torch.complex64).contiguous()
                ideal = ideal_channel.to(self.device, torch.complex64).contiguous()
                stream = torch.cuda.current_stream(self.device).cuda_stream
                _capi.check(_capi.lib().aft_error_sums(C.c_void_p(est.data_ptr()), C.c_void_p(ideal.data_ptr()), est.numel(),
                                                      C.c_void_p(sums.data_ptr()), C.c_void_p(stream)))
                elements += est.numel()
        if elements == 0:
            raise ValueError("empty data loader")
        return float(sums[0].item()) / elements

    def get_test_stats(self, test_dataloaders: List[Tuple[str, DataLoader]], loss_fn=None) -> Dict[int, float]:
        stats = {}
        for name, loader in sorted(test_dataloaders, key=lambda x: int(x[0].split("_")[1])):
            var, val = name.split("_")
            db_error = to_db(self._evaluate_dataloader(loader, loss_fn))
            self.logger.info(f"{var}:{val} Test MSE: {db_error:.4f} dB")
            stats[int(val)] = db_error
        return stats

    def predict_channels(self, test_dataloaders: List[Tuple[str, DataLoader]]) -> Dict[int, Dict]:
        channels = {}
        for name, loader in sorted(test_dataloaders, key=lambda x: int(x[0].split("_")[1])):
            with torch.no_grad():
                est_in, ideal, meta = next(iter(loader))
                est = self._forward_pass(est_in, self.model, meta)
            channels[int(name.split("_")[1])] = {"estimated_channel": est[0], "ideal_channel": ideal[0]}
        return channels


def load_checkpoint(model, checkpoint_path: Union[str, Path]) -> int:
    """Load the ``model_state_dict`` of a reference checkpoint (trainer.py:648-703; keys and shapes are identical, SURVEY
    App. A).  A bare state_dict file is accepted too.  Returns the stored epoch (or -1)."""
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    state = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt
    model.load_state_dict(state)
    return int(ckpt.get("epoch", -1)) if isinstance(ckpt, dict) else -1


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(description="Evaluate a B200 estimator on DS/MDS/SNR-style test folders")
    ap.add_argument("--system_config_path", required=True)
    ap.add_argument("--model_config_path", required=True)
    ap.add_argument("--model_name", choices=sorted(MODEL_REGISTRY), required=True)
    ap.add_argument("--test_set", required=True, help="directory whose sub-directories are named VAR_value")
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--precision", choices=("fp32", "bf16"), default="fp32")
    ap.add_argument("--batch_size", type=int, default=512)
    args = ap.parse_args(argv)
    system_config, model_config = load_config(args.system_config_path, args.model_config_path)
    if str(model_config.device) == "cpu":
        model_config.device = "cuda"
    model = MODEL_REGISTRY[args.model_name](system_config, model_config).eval()
    if args.checkpoint:
        load_checkpoint(model, args.checkpoint)
    if hasattr(model, "precision"):
        model.precision = args.precision
    loaders = get_test_dataloaders(args.test_set, system_config.pilot, args.batch_size, device=model.device)
    stats = ModelEvaluator(model, model.device).get_test_stats(loaders)
    print(json.dumps({str(k): v for k, v in stats.items()}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
