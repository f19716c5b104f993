"""CPU: the drop-in boundary -- config schemas, state_dict contract, error behaviour, C-ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from adafortitran_b200 import (AdaFortiTranEstimator, FortiTranEstimator, ModelConfig, SystemConfig, _capi,
                               load_config)
from tests import util


def test_yaml_configs_load():
    sc, mc = load_config(os.path.join(util.ROOT, "config/system_config.yaml"),
                         os.path.join(util.ROOT, "config/adafortitran.yaml"))
    assert (sc.ofdm.num_scs, sc.ofdm.num_symbols, sc.pilot.num_scs, sc.pilot.num_symbols) == (120, 14, 12, 2)
    assert mc.model_type == "adafortitran" and mc.channel_adaptivity_hidden_sizes == [7, 42, 560]
    assert mc.device == "cpu"
    _, mf = load_config(os.path.join(util.ROOT, "config/system_config.yaml"),
                        os.path.join(util.ROOT, "config/fortitran.yaml"))
    assert mf.model_type == "fortitran" and mf.adaptive_token_length is None


def test_reference_import_paths():
    from src.config import load_config as lc            # noqa: F401
    from src.config.schemas import ModelConfig as MC, SystemConfig as SC  # noqa: F401
    from src.models import AdaFortiTranEstimator as A, FortiTranEstimator as F
    assert A is AdaFortiTranEstimator and F is FortiTranEstimator
    # the evaluation half of the reference tree (trainer.py:259-347 and what it imports)
    from src.config.config_loader import ConfigLoader, load_config as lc2   # noqa: F401
    from src.data import MatDataset, get_test_dataloaders                    # noqa: F401
    from src.main.trainer import MODEL_REGISTRY, ModelEvaluator               # noqa: F401
    from src.models.adafortitran import AdaFortiTranEstimator as A2
    from src.models.fortitran import BaseFortiTranEstimator, FortiTranEstimator as F2   # noqa: F401
    from src.models.linear import LinearEstimator                             # noqa: F401
    from src.utils import concat_complex_channel, extract_values, mse, to_db   # noqa: F401
    assert A2 is A and F2 is F and MODEL_REGISTRY["adafortitran"] is A
    x = torch.complex(torch.arange(6.).reshape(1, 2, 3), -torch.arange(6.).reshape(1, 2, 3))
    assert torch.equal(concat_complex_channel(x), torch.cat((x.real, x.imag), dim=1)) and abs(to_db(100.0) - 20.0) < 1e-12


def test_schema_validation_errors(tmp_path):
    with pytest.raises(ValueError):
        SystemConfig(ofdm=dict(num_scs=12, num_symbols=14), pilot=dict(num_scs=13, num_symbols=2))
    with pytest.raises(ValueError):
        ModelConfig(**{**util.ADA, "adaptive_token_length": None})
    with pytest.raises(ValueError):
        ModelConfig(**{**util.FORTI, "adaptive_token_length": 6})
    with pytest.raises(ValueError):
        ModelConfig(**{**util.FORTI, "unknown_field": 1})
    with pytest.raises(ValueError):
        ModelConfig(**{**util.FORTI, "device": "tpu"})
    with pytest.raises(FileNotFoundError):
        load_config(tmp_path / "nope.yaml", tmp_path / "nope2.yaml")
    bad = tmp_path / "sys.yaml"
    bad.write_text("ofdm: {num_scs: 0, num_symbols: 14}\npilot: {num_scs: 12, num_symbols: 2}\n")
    ok = tmp_path / "m.yaml"
    ok.write_text("model_type: fortitran\npatch_size: [3, 2]\nnum_layers: 6\nmodel_dim: 128\nnum_head: 4\n")
    with pytest.raises(ValueError):
        load_config(bad, ok)


def test_state_dict_contract_matches_reference():
    sd = util.ada_weights()
    m = util.make_model("ada", device="cpu")
    ours = m.state_dict()
    assert list(ours.keys()) == list(sd.keys())
    for k, v in sd.items():
        assert tuple(ours[k].shape) == v.shape, k
        assert ours[k].dtype == torch.float32
    m.load_state_dict(util.to_torch(sd))      # strict
    info = m.get_model_info()
    assert info["total_parameters"] == 987746 and info["transformer_input_dim"] == 12
    f = util.make_model("forti", device="cpu")
    assert f.get_model_info()["total_parameters"] == 913688
    assert not any(k.startswith("channel_adapter") for k in f.state_dict())
    assert f.state_dict()["transformer_encoder.linear_1.weight"].shape == (128, 6)
    s = util.make_model("ada", device="cpu", overrides={"pos_encoding_type": "sinusoidal"})
    assert "transformer_encoder.positional_encoding.pe" in s.state_dict()
    v = util.golden("golden_variants.npz")
    np.testing.assert_allclose(s.state_dict()["transformer_encoder.positional_encoding.pe"][0, :280].numpy(),
                               v["pe_first_rows"], atol=1e-6)


def test_attributes_and_isinstance_dispatch():
    from adafortitran_b200 import BaseFortiTranEstimator
    a = util.make_model("ada", device="cpu")
    f = util.make_model("forti", device="cpu")
    assert isinstance(a, AdaFortiTranEstimator) and not isinstance(f, AdaFortiTranEstimator)
    assert isinstance(a, BaseFortiTranEstimator) and isinstance(a, torch.nn.Module)
    assert a.ofdm_size == (120, 14) and a.pilot_size == (12, 2) and a.patch_length == 6
    assert a.use_channel_adaptation and not f.use_channel_adaptation
    assert a.device == torch.device("cpu") and f.transformer_input_dim == 6
    assert len(list(a.named_parameters())) == len(util.ada_weights())


def test_constructor_errors():
    sc = SystemConfig(**util.SYS)
    mc = ModelConfig(**util.FORTI)          # no adapter fields
    with pytest.raises(ValueError, match="adaptive_token_length"):
        AdaFortiTranEstimator(sc, mc)
    mc2 = ModelConfig(**{**util.ADA, "channel_adaptivity_hidden_sizes": [7, 42]})
    with pytest.raises(ValueError, match="exactly 3"):
        AdaFortiTranEstimator(sc, mc2)


def test_forward_errors_without_gpu_path():
    a = util.make_model("ada", device="cpu")
    x = torch.zeros(2, 12, 2, dtype=torch.cfloat)
    with pytest.raises(ValueError, match="meta_data is required"):
        a(x)
    # no CPU fallback: a CPU-resident model refuses to run
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        a(x, util.meta([0, 0], [50, 50], [200, 200]))
    a.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        a(x, util.meta([0, 0], [50, 50], [200, 200]))


def test_forward_host_validates_like_forward():
    """forward_host used to skip the argument checks: a wrong pilot shape / meta length / output buffer must be a
    Python error, never an out-of-bounds host read or write inside the library."""
    a = util.make_model("ada", device="cpu")
    md = util.meta([0, 0], [50, 50], [200, 200])
    with pytest.raises(ValueError, match="meta_data is required"):
        a.forward_host(torch.zeros(2, 12, 2, dtype=torch.cfloat))
    with pytest.raises(ValueError, match="expected pilot_symbols of shape"):
        a.forward_host(torch.zeros(2, 11, 2, dtype=torch.cfloat), md)
    with pytest.raises(TypeError, match="complex"):
        a.forward_host(torch.zeros(2, 12, 2), md)
    with pytest.raises(ValueError, match="must have 2 elements"):
        a.forward_host(torch.zeros(2, 12, 2, dtype=torch.cfloat), util.meta([0] * 3, [50] * 3, [200] * 3))
    for bad in (torch.zeros(1, 120, 14, dtype=torch.cfloat), torch.zeros(2, 120, 14), torch.zeros(2, 14, 120, dtype=torch.cfloat).transpose(1, 2)):
        with pytest.raises(ValueError, match="out must be a contiguous CPU complex64"):
            a.forward_host(torch.zeros(2, 12, 2, dtype=torch.cfloat), md, out=bad)
    a.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        a.forward_host(torch.zeros(2, 12, 2, dtype=torch.cfloat), md)


def test_weight_staleness_controls():
    a = util.make_model("forti", device="cpu")
    a._packed_key = ("something",)
    a.invalidate_weights()
    assert a._packed_key is None
    a._packed_key = ("something",)
    a.load_state_dict(a.state_dict())          # load_state_dict always forces a repack
    assert a._packed_key is None
    a.weight_check = "bogus"
    with pytest.raises(ValueError, match="weight_check"):
        a._sync_weights()


def test_linear_estimator_follows_its_parameters():
    from adafortitran_b200 import LinearEstimator, ModelConfig, SystemConfig
    m = LinearEstimator(SystemConfig(**util.SYS), ModelConfig(**dict(util.FORTI, model_type="linear", device="cpu"))).eval()
    assert m.device.type == "cpu"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            m(torch.zeros(2, 12, 2))
    m.to(torch.float64)                          # _apply keeps .device in sync with the parameters
    assert m.device == m.linear.weight.device


def test_patch_maps_match_oracle():
    from oracle import aft_oracle as O
    m = util.make_model("forti", device="cpu")
    img = torch.randn(3, 120, 14)
    tok = m.patch_embedder(img)
    np.testing.assert_array_equal(tok.numpy(), O.patchify(img.numpy(), (3, 2)))
    np.testing.assert_array_equal(m.patch_reconstructor(tok).numpy(), img.numpy())


def test_c_abi_library_loads_and_exports_header_symbols():
    from adafortitran_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(util.ROOT, "include", "aft.h")).read()
    declared = set(re.findall(r"AFT_API\s+[\w\s\*]+?\b(aft_\w+)\s*\(", header))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert _capi.lib().aft_abi_version() == _capi.AFT_ABI_VERSION
    # argument validation that needs no GPU
    assert _capi.lib().aft_create(None, None) == _capi.AFT_ERR_INVALID
    assert b"NULL" in _capi.lib().aft_last_error()
    cfg = _capi.AftConfig(num_scs=64, num_symbols=14)
    h = ctypes.c_void_p()
    assert _capi.lib().aft_create(ctypes.byref(cfg), ctypes.byref(h)) == _capi.AFT_ERR_UNSUPPORTED
    assert ctypes.sizeof(_capi.AftConfig) == 17 * 4


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (charter: no CPU fallback through the checker)."""
    pkg = os.path.join(util.ROOT, "adafortitran_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
