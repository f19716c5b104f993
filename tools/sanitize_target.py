"""Small workloads for compute-sanitizer (tools/sanitize.sh): every SIMT kernel family of the library at tiny batch.
usage: python tools/sanitize_target.py fp32|generic|aux|bf16"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafortitran_b200 import _capi, data
from oracle import aft_oracle as O          # synthetic inputs only
from tests import util

what = sys.argv[1]
sd = util.ada_weights()
p, snr, ds, dop = O.synthetic_batch(3, seed=5)
with torch.no_grad():
    if what in ("fp32", "bf16"):
        for kind in ("ada", "forti"):
            m = util.make_model(kind, weights=sd if kind == "ada" else util.forti_weights(sd), precision=what)
            y = m(torch.from_numpy(p), util.meta(snr, ds, dop) if kind == "ada" else None)
            yh = m.forward_host(torch.from_numpy(p), util.meta(snr, ds, dop) if kind == "ada" else None)
            torch.cuda.synchronize()
            assert torch.isfinite(torch.view_as_real(y)).all() and torch.equal(y.cpu(), yh)
    elif what == "generic":
        from tests.test_generic_grid import G, CASES, _model
        for tag in ("a", "f"):
            for prec in ("fp32", "bf16"):
                m = _model(tag, precision=prec)
                md = util.meta(G[tag + "/snr"], G[tag + "/ds"], G[tag + "/dop"]) if CASES[tag]["kind"] == "ada" else None
                y = m(torch.from_numpy(G[tag + "/pilots"]), md)
                torch.cuda.synchronize()
                assert torch.isfinite(torch.view_as_real(y)).all()
    elif what == "aux":
        lib = _capi.lib()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        est = torch.view_as_complex(torch.randn(5, 120, 14, 2, device="cuda"))
        tru = torch.view_as_complex(torch.randn(5, 120, 14, 2, device="cuda"))
        sums = torch.zeros(2, dtype=torch.float64, device="cuda")
        _capi.check(lib.aft_error_sums(C.c_void_p(est.data_ptr()), C.c_void_p(tru.data_ptr()), est.numel(), C.c_void_p(sums.data_ptr()), st))
        grid = torch.zeros(5, 120, 14, dtype=torch.complex64)
        grid[:, ::10, [2, 11]] = torch.view_as_complex(torch.randn(5, 12, 2, 2))
        pil = data.extract_pilots(grid.cuda(), (12, 2))
        from adafortitran_b200 import LinearEstimator, ModelConfig, SystemConfig
        lin = LinearEstimator(SystemConfig(**util.SYS), ModelConfig(**dict(util.FORTI, model_type="linear", device="cuda"))).eval()
        y = lin(torch.randn(5, 12, 2, device="cuda"))
        torch.cuda.synchronize()
        assert pil.shape == (5, 12, 2) and y.shape == (5, 120, 14) and float(sums[1]) > 0
print("sanitize target", what, "ok")
