"""Roofline of the "next"-row kernels (N1 error sums, N2 pilot extraction, N3 LinearEstimator): HBM bound, achieved GB/s
of algorithmic bytes against MEASURED_PEAKS.json.  Inputs larger than L2 (126 MB); CUDA events on the launch stream.
usage: python tools/aux_bench.py > profiles/r01_aux_kernels.json"""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adafortitran_b200 import _capi, data, LinearEstimator, ModelConfig, SystemConfig
from tests import util

def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    peaks = {}
    try: peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception: pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    out = {"peak_hbm_gbs": hbm, "peak_source": src, "kernels": {}}
    dev = "cuda"
    # N2: pilot extraction, B = 65536 grids of 120 x 14 complex64 (881 MB in, 12.6 MB out)
    B = 65536
    grid = torch.zeros(B, 120, 14, dtype=torch.cfloat, device=dev)
    grid[:, 0:120:10, 2] = torch.randn(B, 12, dtype=torch.cfloat, device=dev)
    grid[:, 0:120:10, 11] = torch.randn(B, 12, dtype=torch.cfloat, device=dev)
    pil = torch.empty(B, 12, 2, dtype=torch.cfloat, device=dev); cnt = torch.empty(B, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    f = lambda: _capi.check(_capi.lib().aft_extract_pilots(C.c_void_p(grid.data_ptr()), C.c_void_p(pil.data_ptr()), C.c_void_p(cnt.data_ptr()), B, 1680, 24, st))
    ms = timed(f); by = B * (1680 * 8 + 24 * 8 + 4)
    out["kernels"]["extract_pilots_kernel"] = {"batch": B, "ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6, "frac": by / ms / 1e6 / hbm,
                                               "samples_per_s": B / ms * 1e3}
    assert int((cnt != 24).sum()) == 0
    del grid
    # N3: LinearEstimator, B = 262144 (25 MB in, 1.76 GB out)
    B = 262144
    m = LinearEstimator(SystemConfig(**util.SYS), ModelConfig(**dict(util.FORTI, model_type="linear", device=dev))).eval()
    x = torch.randn(B, 24, device=dev); y = torch.empty(B, 1680, device=dev)
    f = lambda: _capi.check(_capi.lib().aft_linear_forward(C.c_void_p(m.linear.weight.data_ptr()), C.c_void_p(m.linear.bias.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), B, 24, 1680, st))
    ms = timed(f); by = B * (24 + 1680) * 4
    ref = torch.nn.functional.linear(x[:256], m.linear.weight, m.linear.bias)
    out["kernels"]["linear_kernel"] = {"batch": B, "ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6, "frac": by / ms / 1e6 / hbm,
                                       "estimates_per_s": B / ms * 1e3, "max_abs_diff_vs_torch_gpu": float((y[:256] - ref).abs().max())}
    del y
    # N1: error sums over 2 x 110 MB complex64 (B = 8192 estimates x 1680)
    n = 8192 * 1680 * 8
    a = torch.randn(n, dtype=torch.cfloat, device=dev); b = torch.randn(n, dtype=torch.cfloat, device=dev)
    sums = torch.zeros(2, dtype=torch.float64, device=dev)
    f = lambda: _capi.check(_capi.lib().aft_error_sums(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), n, C.c_void_p(sums.data_ptr()), st))
    ms = timed(f); by = n * 16
    out["kernels"]["error_sums_kernel"] = {"elements": n, "ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6, "frac": by / ms / 1e6 / hbm}
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
