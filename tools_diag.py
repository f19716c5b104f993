"""Scratch diagnostics run on the GPU box: stage-by-stage comparison of the fp32 path's workspace buffers with
the oracle (enh plane and final residual stream)."""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import aft_oracle as O
from tests import util

def main():
    sd = util.ada_weights()
    g = util.golden("golden_ada.npz")
    m = util.make_model("ada", weights=sd)
    with torch.no_grad():
        y = m(torch.from_numpy(g["pilots"]), util.meta(g["snr"], g["ds"], g["dop"]))
    torch.cuda.synchronize()
    st = {}
    ref = O.forward(util.oracle_cfg(), sd, g["pilots"], g["snr"], g["ds"], g["dop"], stages=st)
    ws = m._workspace
    off = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
    f = ws[off:].view(torch.float32)
    nseq = 16
    enh = f[: nseq * 1680].cpu().numpy().reshape(8, 2, 1680)
    a = (nseq * 1680 + 63) // 64 * 64
    h = f[a: a + nseq * 280 * 128].cpu().numpy().reshape(8, 2, 280, 128)
    res = {"out": O.normwise_err(y.cpu().numpy(), ref)}
    for part, key in ((0, "re"), (1, "im")):
        res[f"enh_{key}"] = O.normwise_err(enh[:, part], st[key]["conv_enhanced"].reshape(8, 1680))
        res[f"h6_{key}"] = O.normwise_err(h[:, part], st[key]["h6"])
    print(json.dumps(res))

if __name__ == "__main__":
    main()
