"""HBM write-only bandwidth (fill of a buffer much larger than L2) next to the copy figure of MEASURED_PEAKS.json:
the roofline of kernels whose traffic is almost all stores (N3 LinearEstimator: 6,720 B written per 96 B read)."""
import json, torch
n = 1 << 29   # 2 GiB of fp32
y = torch.empty(n, device="cuda"); x = torch.empty(n, device="cuda")
def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
ms_fill = timed(lambda: y.fill_(1.0))
ms_zero = timed(lambda: y.zero_())
ms_copy = timed(lambda: y.copy_(x))
ms_sum = timed(lambda: x.sum())
print(json.dumps({"bytes": 4 * n, "fill_gbs": 4 * n / ms_fill / 1e6, "memset_gbs": 4 * n / ms_zero / 1e6,
                  "copy_gbs_read_plus_write": 8 * n / ms_copy / 1e6, "read_gbs_sum": 4 * n / ms_sum / 1e6}))
