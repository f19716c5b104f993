"""debug: bf16 vs fp32 path, 1 layer, error per patch row (40 rows of 7 tokens)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aft_oracle as O
from tests import util
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sd = util.forti_weights(util.ada_weights())
sd = {k: a for k, a in sd.items() if not any(f"layers.{i}." in k for i in range(L, 6))}
p, *_ = O.synthetic_batch(4, seed=3)
out = {}
for prec in ("fp32", "bf16"):
    m = util.make_model("forti", weights=sd, precision=prec, overrides={"num_layers": L})
    with torch.no_grad():
        out[prec] = m(torch.from_numpy(p), None).cpu().numpy()
d = np.abs(out["bf16"] - out["fp32"])           # [B,120,14]
ref = np.abs(out["fp32"]).mean()
rows = d.reshape(4, 40, 3, 14).mean(axis=(0, 2, 3)) / ref
print("L", L, "finite", np.isfinite(out["bf16"].view(np.float32)).all(), "rel err per patch row:")
print(np.array2string(rows, precision=3, max_line_width=200))
