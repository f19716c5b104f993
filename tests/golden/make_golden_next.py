"""Golden vectors for the "next" rows N2 / N3, generated from the LIVE reference in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_next.py      (needs /root/reference; writes golden_next.npz)

* N3: reference ``LinearEstimator`` (src/models/linear.py) under torch.manual_seed(0): weights, seeded real input, output.
* N2: reference ``MatDataset`` (src/data/dataset.py) run on synthetic .mat files written here with scipy: sparse LS grids
  in, the pilots / truth / metadata it returns out; reference ``extract_values`` on a list of file names (valid and not).
The reference's src.utils imports matplotlib / prettytable, which are not installed: both are stubbed (unused here).
"""
import os, sys, tempfile, types
import numpy as np
import scipy.io as sio
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
for name in ("matplotlib", "matplotlib.pyplot", "prettytable"):
    m = types.ModuleType(name)
    if name == "prettytable":
        m.PrettyTable = object
    sys.modules.setdefault(name, m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

from src.config.schemas import SystemConfig, ModelConfig          # noqa: E402
from src.models.linear import LinearEstimator                     # noqa: E402
from src.utils import extract_values                              # noqa: E402
from src.data.dataset import MatDataset                           # noqa: E402

out = {}
sysc = SystemConfig(ofdm={"num_scs": 120, "num_symbols": 14}, pilot={"num_scs": 12, "num_symbols": 2})
modc = ModelConfig(model_type="linear", device="cpu", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4)
torch.manual_seed(0)
lin = LinearEstimator(sysc, modc).eval()
g = torch.Generator().manual_seed(5)
x = torch.randn(16, 12, 2, generator=g)
with torch.no_grad():
    y = lin(x)
out["lin_weight"] = lin.linear.weight.detach().numpy()
out["lin_bias"] = lin.linear.bias.detach().numpy()
out["lin_x"] = x.numpy()
out["lin_y"] = y.numpy()

# ---- N2: synthetic .mat files -> reference MatDataset
rng = np.random.default_rng(11)
names = ["1_SNR-20_DS-50_DOP-500_N-3_TDL-A.mat", "2_SNR-0_DS-350_DOP-1400_N-3_TDL-B.mat", "17_SNR-30_DS-100_DOP-200_N-2_CDL-C.mat",
         "3_SNR-5_DS-150_DOP-800_N-3_TDL-D.mat"]
tmp = tempfile.mkdtemp()
grids = []
for n in names:
    H = np.zeros((120, 14, 3), dtype=np.complex64)
    H[:, :, 0] = (rng.standard_normal((120, 14)) + 1j * rng.standard_normal((120, 14))).astype(np.complex64)
    ls = np.zeros((120, 14), dtype=np.complex64)
    ls[0:120:10, [2, 11]] = (rng.standard_normal((12, 2)) + 1j * rng.standard_normal((12, 2))).astype(np.complex64)
    H[:, :, 1] = ls
    sio.savemat(os.path.join(tmp, n), {"H": H})
    grids.append(H)
ds = MatDataset(tmp, sysc.pilot)
order = [p.name for p in ds.file_list]
items = [ds[i] for i in range(len(ds))]
out["mat_names"] = np.array(order)
out["mat_H"] = np.stack([grids[names.index(n)] for n in order])
out["mat_pilots"] = np.stack([it[0].numpy() for it in items])
out["mat_truth"] = np.stack([it[1].numpy() for it in items])
out["mat_meta"] = np.stack([np.array([float(it[2][k]) for k in range(5)], dtype=np.float32) for it in items])
out["mat_channel"] = np.array([it[2][5][0] for it in items])

# ---- extract_values on names (valid / invalid)
probe = names + ["x_SNR-20_DS-50_DOP-500_N-3_TDL-A.mat", "1_SNR-20_DS-50_DOP-500_N-3_tdl.mat", "1_SNR-20_DS-50_DOP-500_N-3_TDL-A.mat.bak",
                 "12_SNR-7_DS-1_DOP-2_N-9_A-B-C.mat", "1_SNR--5_DS-50_DOP-500_N-3_TDL-A.mat"]
vals, ok = [], []
for n in probe:
    try:
        v = extract_values(n)
        vals.append([float(t) for t in v[:5]]); ok.append(v[5][0])
    except ValueError:
        vals.append([np.nan] * 5); ok.append("")
out["ev_names"] = np.array(probe)
out["ev_values"] = np.array(vals, dtype=np.float32)
out["ev_channel"] = np.array(ok)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_next.npz"), **out)
print({k: getattr(v, "shape", None) for k, v in out.items()})
