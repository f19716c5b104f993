"""Golden vectors for non-default grids (the shape-generic AFT_FP32 path), from the LIVE reference:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_generic.py      (needs /root/reference)

Two small configurations with weights from the reference constructors under torch.manual_seed(0):
  a: AdaFortiTran, grid 36 x 8, pilots 6 x 2, patch 3 x 2 (48 tokens), 2 layers, gelu, adapter [7, 42, 96]
  f: FortiTran,    grid 24 x 8, pilots 4 x 2, patch 2 x 4 (24 tokens), 3 layers, relu
Stored per configuration: the full state_dict, inputs, and the reference's fp32 forward output.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AFT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from src.config.schemas import ModelConfig, SystemConfig  # noqa: E402
from src.models import AdaFortiTranEstimator, FortiTranEstimator  # noqa: E402

CONFIGS = {
    "a": dict(cls=AdaFortiTranEstimator, sys=dict(ofdm=dict(num_scs=36, num_symbols=8), pilot=dict(num_scs=6, num_symbols=2)),
              model=dict(model_type="adafortitran", patch_size=(3, 2), num_layers=2, model_dim=128, num_head=4, activation="gelu",
                         dropout=0.1, max_seq_len=64, pos_encoding_type="learnable", channel_adaptivity_hidden_sizes=[7, 42, 96],
                         adaptive_token_length=6)),
    "f": dict(cls=FortiTranEstimator, sys=dict(ofdm=dict(num_scs=24, num_symbols=8), pilot=dict(num_scs=4, num_symbols=2)),
              model=dict(model_type="fortitran", patch_size=(2, 4), num_layers=3, model_dim=128, num_head=4, activation="relu",
                         dropout=0.1, max_seq_len=32, pos_encoding_type="learnable")),
}


def main():
    out = {}
    for tag, c in CONFIGS.items():
        torch.manual_seed(0)
        m = c["cls"](SystemConfig(**c["sys"]), ModelConfig(**c["model"])).eval()
        g = torch.Generator().manual_seed(17)
        B = 5
        ps = (c["sys"]["pilot"]["num_scs"], c["sys"]["pilot"]["num_symbols"])
        x = torch.complex(torch.randn(B, *ps, generator=g), torch.randn(B, *ps, generator=g))
        snr = torch.tensor([0., 10., 20., 30., 15.]).reshape(B, 1)
        ds = torch.tensor([50., 150., 250., 350., 100.]).reshape(B, 1)
        dop = torch.tensor([200., 600., 1000., 1400., 800.]).reshape(B, 1)
        meta = (torch.zeros(B, 1), snr, ds, dop, torch.zeros(B, 1), [("syn",) * B])
        with torch.no_grad():
            y = m(x, meta) if tag == "a" else m(x)
        for k, v in m.state_dict().items():
            out[f"{tag}/sd/{k}"] = v.detach().numpy()
        out[f"{tag}/pilots"] = x.numpy()
        out[f"{tag}/snr"], out[f"{tag}/ds"], out[f"{tag}/dop"] = snr.numpy().ravel(), ds.numpy().ravel(), dop.numpy().ravel()
        out[f"{tag}/out"] = y.numpy()
        print(tag, tuple(y.shape), float(y.abs().max()))
    np.savez_compressed(os.path.join(HERE, "golden_generic.npz"), **out)


if __name__ == "__main__":
    main()
