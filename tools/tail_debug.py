"""debug: bf16 vs fp32 path, error per patch row (40 rows of 7 tokens).  usage: tail_debug.py [layers] [forti|ada]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aft_oracle as O
from tests import util
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1
kind = sys.argv[2] if len(sys.argv) > 2 else "forti"
sd = util.ada_weights()
if kind != "ada": sd = util.forti_weights(sd)
sd = {k: a for k, a in sd.items() if not any(f"layers.{i}." in k for i in range(L, 6))}
B = 16
p, snr, ds, dop = O.synthetic_batch(B, seed=7)
out = {}
for prec in ("fp32", "bf16"):
    m = util.make_model(kind, weights=sd, precision=prec, overrides={"num_layers": L})
    with torch.no_grad():
        out[prec] = m(torch.from_numpy(p), util.meta(snr, ds, dop) if kind == "ada" else None).cpu().numpy()
d = np.abs(out["bf16"] - out["fp32"])           # [B,120,14]
ref = np.abs(out["fp32"]).mean()
rows = d.reshape(B, 40, 3, 14).mean(axis=(0, 2, 3)) / ref
mx = d.reshape(B, 40, 3, 14).max(axis=(0, 2, 3)) / np.abs(out["fp32"]).max()
print("L", L, kind, "finite", np.isfinite(out["bf16"].view(np.float32)).all(), "mean rel err per patch row:")
print(np.array2string(rows, precision=3, max_line_width=220))
print("max err / max ref per patch row:")
print(np.array2string(mx, precision=3, max_line_width=220))
print("per-sample max-abs rel:", np.array2string(d.reshape(B, -1).max(axis=1) / np.abs(out["fp32"]).reshape(B, -1).max(axis=1), precision=3, max_line_width=220))
