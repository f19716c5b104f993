"""Rows N2 / N3 of SURVEY.md §8f: input side (file-name metadata, pilot extraction, .mat dataset) and LinearEstimator.
CPU tests pin the oracle and the host logic on golden vectors generated from the live reference
(tests/golden/make_golden_next.py); GPU tests compare the CUDA kernels with the oracle (bit-exact for the index / byte
work of the extraction, 1e-6 normwise for the fp32 linear map) through the C-ABI."""
import os
import tempfile

import numpy as np
import pytest
import torch

from adafortitran_b200 import _capi, data
from oracle import aft_oracle as O
from tests import util

G = util.golden("golden_next.npz")


# ----------------------------------------------------------------------------------------- CPU
def test_oracle_extract_pilots_matches_reference_dataset():
    got = O.extract_pilots(G["mat_H"][..., 1], (12, 2))
    assert got.dtype == np.complex64 and np.array_equal(got, G["mat_pilots"])          # bit-exact


def test_oracle_extract_pilots_rejects_wrong_count():
    ls = G["mat_H"][..., 1].copy()
    ls[0, 0, 2] = 0                                                                     # one pilot lost
    with pytest.raises(ValueError, match="Expected 24 pilot values, got 23"):
        O.extract_pilots(ls, (12, 2))


def test_oracle_linear_matches_reference():
    y = O.linear_estimator(G["lin_weight"], G["lin_bias"], G["lin_x"], (120, 14))
    assert O.normwise_err(y, G["lin_y"]) <= 1e-6


def test_extract_values_matches_reference():
    for name, vals, ch in zip(G["ev_names"], G["ev_values"], G["ev_channel"]):
        if ch == "":
            with pytest.raises(ValueError, match="Cannot extract file information"):
                data.extract_values(str(name))
        else:
            got = data.extract_values(str(name))
            assert [float(t) for t in got[:5]] == [float(v) for v in vals] and got[5] == [str(ch)]
            assert all(t.dtype == torch.float32 and tuple(t.shape) == (1,) for t in got[:5])


def test_mat_dataset_matches_reference(tmp_path):
    import scipy.io as sio
    for name, H in zip(G["mat_names"], G["mat_H"]):
        sio.savemat(tmp_path / str(name), {"H": H})
    ds = data.MatDataset(tmp_path, (12, 2))
    assert len(ds) == 4
    by_name = {p.name: i for i, p in enumerate(ds.file_list)}
    for j, name in enumerate(G["mat_names"]):
        h_est, h_ideal, meta = ds[by_name[str(name)]]
        assert h_est.dtype == torch.complex64 and np.array_equal(h_est.numpy(), G["mat_pilots"][j])
        assert np.array_equal(h_ideal.numpy(), G["mat_truth"][j])
        assert [float(t) for t in meta[:5]] == [float(v) for v in G["mat_meta"][j]] and meta[5] == [str(G["mat_channel"][j])]
    with pytest.raises(IndexError):
        ds[4]
    with pytest.raises(FileNotFoundError):
        data.MatDataset(tmp_path / "missing", (12, 2))
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(ValueError, match="No .mat files"):
        data.MatDataset(empty, (12, 2))
    sio.savemat(tmp_path / "9_SNR-1_DS-2_DOP-3_N-4_TDL-A.mat", {"G": np.zeros(3)})
    bad = data.MatDataset(tmp_path, (12, 2))
    with pytest.raises(ValueError, match="Invalid .mat file format"):
        bad[[p.name for p in bad.file_list].index("9_SNR-1_DS-2_DOP-3_N-4_TDL-A.mat")]


def test_linear_estimator_contract_on_cpu():
    from adafortitran_b200 import LinearEstimator, ModelConfig, SystemConfig
    from src.models import LinearEstimator as Shim
    assert Shim is LinearEstimator
    cfg = dict(util.FORTI, model_type="linear", device="cpu")
    m = LinearEstimator(SystemConfig(**util.SYS), ModelConfig(**cfg)).eval()
    assert sorted(m.state_dict()) == ["linear.bias", "linear.weight"] and tuple(m.linear.weight.shape) == (1680, 24)
    assert m.ofdm_size == (120, 14) and m.pilot_size == (12, 2) and "LinearEstimator(" in repr(m)
    with pytest.raises(ValueError, match="Expected input shape"):
        m(torch.zeros(2, 24))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            m(torch.zeros(2, 12, 2))
    with pytest.raises(RuntimeError, match="complex"):
        m(torch.zeros(2, 12, 2, dtype=torch.cfloat))


# ----------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_extract_pilots_bit_exact_and_edges():
    ls = torch.from_numpy(G["mat_H"][..., 1].copy()).cuda()
    got = data.extract_pilots(ls, (12, 2))
    assert np.array_equal(got.cpu().numpy(), G["mat_pilots"])
    # random sparse grids: random positions (not a regular comb), negative zero and NaN entries, ragged batch sizes
    rng = np.random.default_rng(3)
    for batch in (1, 7, 33, 1000):
        grid = np.zeros((batch, 120 * 14), dtype=np.complex64)
        for b in range(batch):
            pos = rng.choice(1680, size=24, replace=False)
            grid[b, pos] = (rng.standard_normal(24) + 1j * rng.standard_normal(24)).astype(np.complex64)
        grid[0, np.flatnonzero(grid[0] == 0)[0]] = complex(-0.0, 0.0)          # -0.0 + 0j is zero for the mask
        if batch > 1:
            nzpos = np.flatnonzero(grid[1])[3]
            grid[1, nzpos] = complex(np.nan, 0.0)                               # NaN != 0: stays a pilot
        grid = grid.reshape(batch, 120, 14)
        want = O.extract_pilots(grid, (12, 2))
        got = data.extract_pilots(torch.from_numpy(grid).cuda(), (12, 2)).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))        # bit pattern, NaN included
    # wrong counts: too few / too many / empty grid
    for delta in (-1, +1):
        grid = G["mat_H"][..., 1].copy()
        if delta < 0:
            grid[2, 0, 2] = 0
        else:
            grid[2, 5, 5] = 1 + 1j
        with pytest.raises(ValueError, match=rf"Expected 24 pilot values, got {24 + delta} \(sample 2\)"):
            data.extract_pilots(torch.from_numpy(grid).cuda(), (12, 2))
    with pytest.raises(ValueError, match="got 0"):
        data.extract_pilots(torch.zeros(3, 120, 14, dtype=torch.cfloat, device="cuda"), (12, 2))
    assert tuple(data.extract_pilots(torch.zeros(0, 120, 14, dtype=torch.cfloat, device="cuda"), (12, 2)).shape) == (0, 12, 2)
    # other grid / pilot sizes (cells not a multiple of 32)
    grid = np.zeros((5, 7 * 9), dtype=np.complex64)
    for b in range(5):
        grid[b, rng.choice(63, size=6, replace=False)] = (rng.standard_normal(6) + 1j).astype(np.complex64)
    grid = grid.reshape(5, 7, 9)
    assert np.array_equal(data.extract_pilots(torch.from_numpy(grid).cuda(), (3, 2)).cpu().numpy(), O.extract_pilots(grid, (3, 2)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        data.extract_pilots(torch.zeros(1, 120, 14, dtype=torch.cfloat), (12, 2))


@pytest.mark.gpu
def test_linear_estimator_parity():
    from adafortitran_b200 import LinearEstimator, ModelConfig, SystemConfig
    cfg = dict(util.FORTI, model_type="linear", device="cuda")
    m = LinearEstimator(SystemConfig(**util.SYS), ModelConfig(**cfg)).eval()
    m.load_state_dict({"linear.weight": torch.from_numpy(G["lin_weight"]), "linear.bias": torch.from_numpy(G["lin_bias"])})
    with torch.no_grad():
        y = m(torch.from_numpy(G["lin_x"]))
    assert y.dtype == torch.float32 and tuple(y.shape) == (16, 120, 14) and y.device.type == "cuda"
    assert O.normwise_err(y.cpu().numpy(), G["lin_y"]) <= 1e-6                  # fp32 gate (sums of 24 products)
    # ragged / large batches against the fp64 oracle; linearity as the size-independent property
    rng = np.random.default_rng(8)
    for batch in (1, 9, 4099):
        x = rng.standard_normal((batch, 12, 2)).astype(np.float32)
        with torch.no_grad():
            y = m(torch.from_numpy(x)).cpu().numpy()
        assert O.normwise_err(y, O.linear_estimator(G["lin_weight"], G["lin_bias"], x, (120, 14))) <= 1e-6
    x1, x2 = (torch.from_numpy(rng.standard_normal((64, 12, 2)).astype(np.float32)) for _ in range(2))
    with torch.no_grad():
        lhs = m(x1 + x2) + m(torch.zeros(64, 12, 2))
        rhs = m(x1) + m(x2)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) <= 1e-6
    with torch.no_grad():
        assert tuple(m(torch.zeros(0, 12, 2)).shape) == (0, 120, 14)
    # other shapes go through the generic kernel of the same entry point
    import ctypes as C
    for in_dim, out_dim, batch in ((7, 33, 19), (24, 3000, 5), (100, 64, 130)):
        w = rng.standard_normal((out_dim, in_dim)).astype(np.float32)
        b = rng.standard_normal(out_dim).astype(np.float32)
        x = rng.standard_normal((batch, in_dim)).astype(np.float32)
        tw, tb, tx = (torch.from_numpy(a).cuda() for a in (w, b, x))
        ty = torch.empty(batch, out_dim, device="cuda")
        _capi.check(_capi.lib().aft_linear_forward(C.c_void_p(tw.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(tx.data_ptr()),
                                                   C.c_void_p(ty.data_ptr()), batch, in_dim, out_dim,
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        want = x.astype(np.float64) @ w.astype(np.float64).T + b
        assert O.normwise_err(ty.cpu().numpy(), want) <= 2e-6


@pytest.mark.gpu
def test_collate_on_device(tmp_path):
    import scipy.io as sio
    for name, H in zip(G["mat_names"], G["mat_H"]):
        sio.savemat(tmp_path / str(name), {"H": H})
    ds = data.MatDataset(tmp_path, (12, 2), raw=True)
    items = [ds[i] for i in range(len(ds))]
    pilots, truth, meta = data.collate_on_device(items, (12, 2))
    order = [list(G["mat_names"]).index(p.name) for p in ds.file_list]
    assert np.array_equal(pilots.cpu().numpy(), G["mat_pilots"][order]) and np.array_equal(truth.cpu().numpy(), G["mat_truth"][order])
    assert all(tuple(t.shape) == (4, 1) for t in meta[:5]) and meta[5] == [tuple(str(G["mat_channel"][i]) for i in order)]
    assert np.array_equal(torch.cat(meta[:5], dim=1).numpy(), G["mat_meta"][order])
    # the collated batch drives the estimator exactly like the reference loader's batch
    m = util.make_model("ada", weights=util.ada_weights(), precision="fp32")
    with torch.no_grad():
        y = m(pilots, meta)
    assert tuple(y.shape) == (4, 120, 14) and bool(torch.isfinite(torch.view_as_real(y)).all())


# ----------------------------------------------------------------------------------------- N4: evaluator
EVAL_SETS = {"DS_50": [f"{i}_SNR-20_DS-50_DOP-500_N-3_TDL-A.mat" for i in range(1, 8)],
             "DS_200": [f"{i}_SNR-10_DS-200_DOP-900_N-3_TDL-B.mat" for i in range(1, 6)]}


def _write_test_sets(root, names_by_dir, seed=21):
    import scipy.io as sio
    rng = np.random.default_rng(seed)
    for sub, names in names_by_dir.items():
        (root / sub).mkdir(parents=True)
        for n in names:
            H = np.zeros((120, 14, 3), dtype=np.complex64)
            H[:, :, 0] = (rng.standard_normal((120, 14)) + 1j * rng.standard_normal((120, 14))).astype(np.complex64) * 0.7
            ls = np.zeros((120, 14), dtype=np.complex64)
            ls[0:120:10, [2, 11]] = H[0:120:10, :, 0][:, [2, 11]] + 0.05 * (rng.standard_normal((12, 2)) + 1j * rng.standard_normal((12, 2)))
            H[:, :, 1] = ls
            sio.savemat(root / sub / n, {"H": H})


@pytest.mark.gpu
def test_evaluator_matches_reference_formula(tmp_path):
    """Folder sweep through the B200 ModelEvaluator vs the reference formula evaluated with the CPU port of the model:
    sum_b 2 * MSELoss(cat(re, im)) * B_b / sum_b B_b, to_db (trainer.py:328-347)."""
    from adafortitran_b200 import evaluate
    from adafortitran_b200.config import PilotParams
    from oracle.torch_port import TorchPort
    sets = EVAL_SETS
    _write_test_sets(tmp_path, sets)
    sd = util.ada_weights()
    model = util.make_model("ada", weights=sd, precision="fp32")
    loaders = evaluate.get_test_dataloaders(tmp_path, PilotParams(num_scs=12, num_symbols=2), batch_size=3)
    assert sorted(n for n, _ in loaders) == ["DS_200", "DS_50"]
    stats = evaluate.ModelEvaluator(model, model.device).get_test_stats(loaders)
    assert list(stats) == [50, 200]
    # the reference's OWN ModelEvaluator.get_test_stats on the same files (tests/golden/make_golden_evaluator.py, live reference)
    ge = util.golden("golden_evaluator.npz")
    assert int(ge["batch_size"]) == 3 and list(ge["keys"]) == [50, 200]
    for k, want in zip(ge["keys"], ge["mse_db"]):
        assert abs(stats[int(k)] - float(want)) <= 2e-3, (k, stats, want)
    # reference side: CPU port of the model + the reference's metric, batch by batch
    port = TorchPort(sd, adaptive=True).eval()
    for sub, names in sets.items():
        ds = data.MatDataset(tmp_path / sub, (12, 2))
        total, n = 0.0, 0
        for i0 in range(0, len(ds), 3):
            items = [ds[i] for i in range(i0, min(i0 + 3, len(ds)))]
            x = torch.stack([it[0] for it in items]); h = torch.stack([it[1] for it in items])
            meta = [torch.stack([it[2][k] for it in items]) for k in range(5)]
            with torch.no_grad():
                y = port(x, meta[1].reshape(-1), meta[2].reshape(-1), meta[3].reshape(-1))
            cat = lambda t: torch.cat((t.real, t.imag), dim=1)
            loss = torch.nn.functional.mse_loss(cat(y), cat(h))
            total += 2 * loss.item() * len(items); n += len(items)
        want = 10 * np.log10(total / n)
        assert abs(stats[int(sub.split("_")[1])] - want) <= 2e-3, (sub, stats, want)
    # checkpoint in the reference's format
    ck = tmp_path / "checkpoint_epoch_3.pt"
    torch.save({"epoch": 3, "model_state_dict": util.to_torch(sd)}, ck)
    fresh = util.make_model("ada", precision="fp32")
    assert evaluate.load_checkpoint(fresh, ck) == 3
    again = evaluate.ModelEvaluator(fresh, fresh.device).get_test_stats(evaluate.get_test_dataloaders(tmp_path, PilotParams(num_scs=12, num_symbols=2), 4))
    assert all(abs(again[k] - stats[k]) <= 1e-6 for k in stats)
    chans = evaluate.ModelEvaluator(fresh, fresh.device).predict_channels(loaders)
    assert sorted(chans) == [50, 200] and tuple(chans[50]["estimated_channel"].shape) == (120, 14)
