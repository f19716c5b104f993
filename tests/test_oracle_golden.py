"""CPU: the oracle restatement against the golden vectors generated from the live reference
(tests/golden/make_golden.py).  This is what pins parity (the reference has no vectors)."""
import numpy as np
import pytest

from oracle import aft_oracle as O
from tests import util


@pytest.fixture(scope="module")
def sd():
    return util.ada_weights()


def test_ada_end_to_end_fp64_and_fp32(sd):
    g = util.golden("golden_ada.npz")
    y64 = O.forward(util.oracle_cfg(), sd, g["pilots"], g["snr"], g["ds"], g["dop"], dtype=np.float64)
    assert O.normwise_err(y64, g["out64"]) < 1e-12          # vs the reference run in .double()
    assert O.normwise_err(y64, g["out"]) < 5e-5             # vs the reference fp32 forward (its own rounding)
    y32 = O.forward(util.oracle_cfg(), sd, g["pilots"], g["snr"], g["ds"], g["dop"], dtype=np.float32)
    assert y32.dtype == np.complex64
    assert O.normwise_err(y32, g["out"]) < 5e-5


def test_ada_stages(sd):
    g = util.golden("golden_ada.npz")
    st = {}
    O.forward(util.oracle_cfg(), sd, g["pilots"][:2], g["snr"][:2], g["ds"][:2], g["dop"][:2], stages=st)
    for name in ("upsampled", "conv_enhanced", "tokens", "h0", "h1", "h6", "tok_out", "combined"):
        ref = g["st_" + name]
        got = st["re"][name].reshape(ref.shape)
        assert O.normwise_err(got, ref) < 5e-5, name
    # adapter layout: [snr0, snr1, ds0, ds1, dop0, dop1] per token
    assert O.normwise_err(st["re"]["tokens"][..., 6:], g["st_adapter"]) < 1e-5


def test_fortitran(sd):
    g = util.golden("golden_forti.npz")
    y = O.forward(util.oracle_cfg("forti"), util.forti_weights(sd), g["pilots"])
    assert O.normwise_err(y, g["out"]) < 5e-5


def test_variants(sd):
    g = util.golden("golden_ada.npz")
    v = util.golden("golden_variants.npz")
    args = (g["pilots"][:4], g["snr"][:4], g["ds"][:4], g["dop"][:4])
    y = O.forward(util.oracle_cfg(activation="relu"), sd, *args)
    assert O.normwise_err(y, v["out_relu"]) < 5e-5
    y = O.forward(util.oracle_cfg(num_layers=2), sd, *args)
    assert O.normwise_err(y, v["out_layers2"]) < 5e-5
    s2 = {k: a for k, a in sd.items() if "position_embeddings" not in k}
    pe = np.zeros((1, 512, 128), dtype=np.float32)
    pe[0, :280] = v["pe_first_rows"]
    s2["transformer_encoder.positional_encoding.pe"] = pe
    y = O.forward(util.oracle_cfg(), s2, *args)
    assert O.normwise_err(y, v["out_sinusoidal"]) < 5e-5


def test_sweep_conditions(sd):
    g = util.golden("golden_sweep.npz")
    for i in (0, 6, 7, 13, 14, 20):     # first/last of each swept axis
        s, d, f = g["conds"][i]
        y = O.forward(util.oracle_cfg(), sd, g["pilots"][i], [s] * 4, [d] * 4, [f] * 4)
        assert O.normwise_err(y, g["out"][i]) < 5e-5, i


def test_missing_meta_raises(sd):
    g = util.golden("golden_ada.npz")
    with pytest.raises(ValueError):
        O.forward(util.oracle_cfg(), sd, g["pilots"])


def test_patchify_roundtrip_and_order():
    img = np.arange(2 * 120 * 14, dtype=np.float64).reshape(2, 120, 14)
    tok = O.patchify(img, (3, 2))
    assert tok.shape == (2, 280, 6)
    assert tok[0, 0].tolist() == [0, 1, 14, 15, 28, 29]      # SURVEY.md Appendix D, probe P3
    assert np.array_equal(O.unpatchify(tok, (120, 14), (3, 2)), img)


def test_metric_definitions():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((4, 120, 14)) + 1j * rng.standard_normal((4, 120, 14))
    b = rng.standard_normal((4, 120, 14)) + 1j * rng.standard_normal((4, 120, 14))
    # reference: 2 * MSELoss(cat(re,im)) == mean |a-b|^2
    cat = lambda z: np.concatenate([z.real, z.imag], axis=1)
    ref = 10 * np.log10(2 * np.mean((cat(a) - cat(b)) ** 2))
    assert abs(O.mse_db_reference(a, b) - ref) < 1e-12


def test_oracle_on_trained_weights():
    """The pseudo-trained fixture (tests/golden/make_golden_trained.py: the live reference after 400 Adam steps): the numpy
    oracle reproduces the reference's recorded outputs, and the fixture is informative (NMSE well below 0 dB)."""
    g = util.golden("golden_trained.npz")
    sd = {k[3:]: g[k] for k in g if k.startswith("sd/")}
    n = int(g["samples"])
    assert float(g["nmse_db"].max()) < -2.5 and float(g["nmse_db"].min()) < -10.0
    for i in (0, 7, 20):
        s, d, f = (float(v) for v in g["conds"][i])
        p, h = O.synthetic_channel(int(g["gen_batch"]), s, d, f, seed=int(g["seed0"]) + i)
        p, h = p[:n], h[:n]
        y = O.forward(util.oracle_cfg("ada"), sd, p, np.full(n, s, np.float32), np.full(n, d, np.float32), np.full(n, f, np.float32),
                      dtype=np.float32)
        assert O.normwise_err(y, g["out"][i]) <= 5e-6
        assert abs(O.nmse_db(g["out"][i], h) - float(g["nmse_db"][i])) < 1e-3
