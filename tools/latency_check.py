"""Small-batch latency of the forward: B = 1 and B = 64 (BASELINE configs[0] batch) for both models and both
precisions.  Per call: device time (CUDA events around model(...)) and host wall time including the synchronisation,
medians over N calls after warm-up.  One JSON line per case."""
import json, os, statistics, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aft_oracle as O          # synthetic inputs only
from tests import util

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
sd = util.ada_weights()
for kind in ("ada", "forti"):
    for prec in ("bf16", "fp32"):
        m = util.make_model(kind, weights=sd if kind == "ada" else util.forti_weights(sd), precision=prec)
        for B in (1, 64):
            p, snr, ds, dop = O.synthetic_batch(B, seed=3)
            x = torch.from_numpy(p).cuda()
            md = tuple(t.cuda() if torch.is_tensor(t) else t for t in util.meta(snr, ds, dop)) if kind == "ada" else None
            with torch.no_grad():
                for _ in range(5):
                    m(x, md)
                torch.cuda.synchronize()
                dev, wall = [], []
                for _ in range(N):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0 = time.perf_counter()
                    e0.record(); y = m(x, md); e1.record()
                    torch.cuda.synchronize()
                    wall.append((time.perf_counter() - t0) * 1e3); dev.append(e0.elapsed_time(e1))
            print(json.dumps({"model": kind, "precision": prec, "batch": B, "device_ms_median": statistics.median(dev),
                              "wall_ms_median": statistics.median(wall), "device_ms_min": min(dev), "calls": N,
                              "estimates_per_s_at_this_batch": B / statistics.median(wall) * 1e3}), flush=True)
