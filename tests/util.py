"""Shared helpers for the test-suite (checker side: may import ``oracle``)."""
import dataclasses
import os

import numpy as np
import torch

from oracle import aft_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SYS = dict(ofdm=dict(num_scs=120, num_symbols=14), pilot=dict(num_scs=12, num_symbols=2))
ADA = dict(model_type="adafortitran", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4,
           activation="gelu", dropout=0.1, max_seq_len=512, pos_encoding_type="learnable",
           channel_adaptivity_hidden_sizes=[7, 42, 560], adaptive_token_length=6)
FORTI = {k: v for k, v in ADA.items() if k not in ("channel_adaptivity_hidden_sizes", "adaptive_token_length")}
FORTI["model_type"] = "fortitran"


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def ada_weights():
    return golden("weights_ada_seed0.npz")


def forti_weights(sd):
    out = {k: v for k, v in sd.items() if not k.startswith("channel_adapter.")}
    out["transformer_encoder.linear_1.weight"] = np.ascontiguousarray(sd["transformer_encoder.linear_1.weight"][:, :6])
    return out


def to_torch(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def meta(snr, ds, dop):
    """The reference 6-tuple (file_no, snr, delay_spread, max_dop_shift, pilot_freq, channel_type)."""
    b = len(snr)
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32)).reshape(b, 1)
    return (torch.zeros(b, 1), t(snr), t(ds), t(dop), torch.zeros(b, 1), [("syn",) * b])


def make_model(kind="ada", device="cuda", precision="fp32", overrides=None, weights=None):
    from adafortitran_b200 import AdaFortiTranEstimator, FortiTranEstimator, ModelConfig, SystemConfig
    cfg = dict(ADA if kind == "ada" else FORTI)
    cfg.update(overrides or {})
    cfg["device"] = device
    cls = AdaFortiTranEstimator if kind == "ada" else FortiTranEstimator
    model = cls(SystemConfig(**SYS), ModelConfig(**cfg)).eval()
    if weights is not None:
        model.load_state_dict(to_torch(weights))
    model.precision = precision
    return model


def oracle_cfg(kind="ada", **kw):
    return dataclasses.replace(O.OracleConfig(), adaptive=(kind == "ada"), **kw)
