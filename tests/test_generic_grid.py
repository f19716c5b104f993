"""Grids other than the reference default (SURVEY.md §8d config 5 and friends): served by the shape-generic AFT_FP32
kernels (csrc/generic_f32.cu) and, with AFT_BF16, by the long-sequence tensor-core encoder (csrc/tc_long.cu).  Golden vectors come from the live reference at two small non-default configurations
(tests/golden/make_golden_generic.py); BASELINE config 5 itself (3276 x 14 grid, 7644 tokens, 154 M parameters) is
checked against the numpy oracle with weights from the model's own seeded initialisation."""
import numpy as np
import pytest
import torch

from oracle import aft_oracle as O
from tests import util

G = util.golden("golden_generic.npz")
CASES = {
    "a": dict(kind="ada", sys=dict(ofdm=dict(num_scs=36, num_symbols=8), pilot=dict(num_scs=6, num_symbols=2)),
              model=dict(model_type="adafortitran", patch_size=(3, 2), num_layers=2, model_dim=128, num_head=4, activation="gelu",
                         max_seq_len=64, pos_encoding_type="learnable", channel_adaptivity_hidden_sizes=[7, 42, 96], adaptive_token_length=6),
              oracle=O.OracleConfig(num_scs=36, num_symbols=8, pilot_scs=6, pilot_symbols=2, patch=(3, 2), num_layers=2, activation="gelu", adaptive=True)),
    "f": dict(kind="forti", sys=dict(ofdm=dict(num_scs=24, num_symbols=8), pilot=dict(num_scs=4, num_symbols=2)),
              model=dict(model_type="fortitran", patch_size=(2, 4), num_layers=3, model_dim=128, num_head=4, activation="relu",
                         max_seq_len=32, pos_encoding_type="learnable"),
              oracle=O.OracleConfig(num_scs=24, num_symbols=8, pilot_scs=4, pilot_symbols=2, patch=(2, 4), num_layers=3, activation="relu", adaptive=False)),
}


def _sd(tag):
    return {k[len(tag) + 4:]: G[k] for k in G if k.startswith(tag + "/sd/")}


def _model(tag, precision="fp32"):
    from adafortitran_b200 import AdaFortiTranEstimator, FortiTranEstimator, ModelConfig, SystemConfig
    c = CASES[tag]
    cls = AdaFortiTranEstimator if c["kind"] == "ada" else FortiTranEstimator
    m = cls(SystemConfig(**c["sys"]), ModelConfig(**dict(c["model"], device="cuda"))).eval()
    m.load_state_dict(util.to_torch(_sd(tag)))
    m.precision = precision
    return m


@pytest.mark.parametrize("tag", ["a", "f"])
def test_oracle_matches_reference_on_other_grids(tag):
    c = CASES[tag]
    y = O.forward(c["oracle"], _sd(tag), G[tag + "/pilots"], G[tag + "/snr"], G[tag + "/ds"], G[tag + "/dop"])
    assert O.normwise_err(y, G[tag + "/out"]) <= 5e-6


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "f"])
def test_fp32_parity_on_other_grids(tag):
    m = _model(tag)
    md = util.meta(G[tag + "/snr"], G[tag + "/ds"], G[tag + "/dop"]) if CASES[tag]["kind"] == "ada" else None
    with torch.no_grad():
        y = m(torch.from_numpy(G[tag + "/pilots"]), md)
    assert tuple(y.shape) == G[tag + "/out"].shape and y.dtype == torch.complex64
    assert O.normwise_err(y.cpu().numpy(), G[tag + "/out"]) <= 1e-4
    # ragged batches / empty batch / host entry point
    with torch.no_grad():
        y1 = m(torch.from_numpy(G[tag + "/pilots"][:1]), tuple(t[:1] if torch.is_tensor(t) else t for t in md) if md else None)
        assert O.normwise_err(y1.cpu().numpy(), G[tag + "/out"][:1]) <= 1e-4
        yh = m.forward_host(torch.from_numpy(G[tag + "/pilots"]), md)
    assert O.normwise_err(yh.numpy(), G[tag + "/out"]) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("tag,gate_db", [("a", -34.0), ("f", -40.0)])
def test_bf16_parity_on_other_grids(tag, gate_db):
    """AFT_BF16 on a non-default grid: the tensor-core encoder over row-tile operand images (csrc/tc_long.cu: flattened-token
    tcgen05 GEMMs + streaming attention).  Gate: output-relative error power vs the reference's recorded output; looser for
    the adaptive model, whose token features carry raw metadata values."""
    m = _model(tag, precision="bf16")
    md = util.meta(G[tag + "/snr"], G[tag + "/ds"], G[tag + "/dop"]) if CASES[tag]["kind"] == "ada" else None
    with torch.no_grad():
        y = m(torch.from_numpy(G[tag + "/pilots"]), md).cpu().numpy()
        y32 = _model(tag)(torch.from_numpy(G[tag + "/pilots"]), md).cpu().numpy()
    assert np.isfinite(y.view(np.float32)).all()
    assert O.rel_err_db(y, G[tag + "/out"]) <= gate_db
    assert O.rel_err_db(y, y32) <= gate_db


@pytest.mark.gpu
def test_baseline_config5_3276x14():
    """BASELINE.json config 5: AdaFortiTran at a 3276 x 14 grid (7644 tokens), dense pilot comb 1638 x 2.
    One sample; weights from the model's own initialisation under a fixed seed; numpy fp64 oracle as the reference."""
    from adafortitran_b200 import AdaFortiTranEstimator, ModelConfig, SystemConfig
    sysc = SystemConfig(ofdm=dict(num_scs=3276, num_symbols=14), pilot=dict(num_scs=1638, num_symbols=2))
    modc = ModelConfig(model_type="adafortitran", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4, activation="gelu",
                       max_seq_len=7644, pos_encoding_type="learnable", channel_adaptivity_hidden_sizes=[7, 42, 15288],
                       adaptive_token_length=6, device="cuda")
    torch.manual_seed(0)
    m = AdaFortiTranEstimator(sysc, modc).eval()
    g = torch.Generator().manual_seed(3)
    x = torch.complex(torch.randn(1, 1638, 2, generator=g), torch.randn(1, 1638, 2, generator=g))
    snr, ds, dop = np.array([20.0], np.float32), np.array([50.0], np.float32), np.array([500.0], np.float32)
    with torch.no_grad():
        y = m(x, util.meta(snr, ds, dop)).cpu().numpy()
    assert y.shape == (1, 3276, 14) and np.isfinite(y.view(np.float32)).all()
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    cfg = O.OracleConfig(num_scs=3276, num_symbols=14, pilot_scs=1638, pilot_symbols=2, patch=(3, 2), num_layers=6, activation="gelu", adaptive=True)
    ref = O.forward(cfg, sd, x.numpy(), snr, ds, dop, dtype=np.float32)
    assert O.normwise_err(y, ref) <= 1e-4
    # the same sample through the long-sequence tensor-core path (60 row tiles of 128 tokens, streaming attention)
    m.precision = "bf16"
    with torch.no_grad():
        yb = m(x, util.meta(snr, ds, dop)).cpu().numpy()
    assert np.isfinite(yb.view(np.float32)).all()
    assert O.rel_err_db(yb, ref) <= -34.0
