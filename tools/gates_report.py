"""Measured values behind the bf16 gates of tests/test_gpu_bf16.py and __graft_entry__.smoke() (one JSON line)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aft_oracle as O
from tests import util
from tests.test_gpu_bf16 import run

sd = util.ada_weights()
g = util.golden("golden_ada.npz"); v = util.golden("golden_variants.npz")
res = {}
m = util.make_model("ada", weights=sd, precision="bf16")
res["ada_vs_out64_db"] = O.rel_err_db(run(m, g["pilots"], g["snr"], g["ds"], g["dop"]), g["out64"])
args = (g["pilots"][:4], g["snr"][:4], g["ds"][:4], g["dop"][:4])
m = util.make_model("ada", weights=sd, precision="bf16", overrides={"activation": "relu"})
res["relu_db"] = O.rel_err_db(run(m, *args), v["out_relu"])
m = util.make_model("ada", precision="bf16", overrides={"num_layers": 2},
                    weights={k: a for k, a in sd.items() if not any(f"layers.{i}." in k for i in range(2, 6))})
res["layers2_db"] = O.rel_err_db(run(m, *args), v["out_layers2"])
m = util.make_model("ada", precision="bf16", overrides={"pos_encoding_type": "sinusoidal"})
s2 = {k: a for k, a in sd.items() if "position_embeddings" not in k}
s2["transformer_encoder.positional_encoding.pe"] = m.state_dict()["transformer_encoder.positional_encoding.pe"].cpu().numpy()
m.load_state_dict(util.to_torch(s2))
res["sinusoidal_db"] = O.rel_err_db(run(m, *args), v["out_sinusoidal"])
pilots, snr, ds, dop = O.synthetic_batch(4, seed=7)
ref = O.forward(util.oracle_cfg(), sd, pilots, snr, ds, dop, dtype=np.float64)
m = util.make_model("ada", weights=sd, precision="bf16")
res["smoke_bf16_db"] = O.rel_err_db(run(m, pilots, snr, ds, dop), ref)
gs = util.golden("golden_sweep.npz")
print(json.dumps({k: round(float(x), 2) for k, x in res.items()}))
