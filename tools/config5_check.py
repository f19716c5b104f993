"""BASELINE config 5 (AdaFortiTran at 3276 x 14, 7644 tokens) through the shape-generic AFT_FP32 path: parity vs the numpy
oracle and a timing of the forward (one JSON line)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafortitran_b200 import AdaFortiTranEstimator, ModelConfig, SystemConfig
from oracle import aft_oracle as O
from tests import util
sysc = SystemConfig(ofdm=dict(num_scs=3276, num_symbols=14), pilot=dict(num_scs=1638, num_symbols=2))
modc = ModelConfig(model_type="adafortitran", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4, activation="gelu",
                   max_seq_len=7644, pos_encoding_type="learnable", channel_adaptivity_hidden_sizes=[7, 42, 15288],
                   adaptive_token_length=6, device="cuda")
torch.manual_seed(0)
m = AdaFortiTranEstimator(sysc, modc).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m.precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
g = torch.Generator().manual_seed(3)
x = torch.complex(torch.randn(B, 1638, 2, generator=g), torch.randn(B, 1638, 2, generator=g))
snr = np.full(B, 20.0, np.float32); ds = np.full(B, 50.0, np.float32); dop = np.full(B, 500.0, np.float32)
md = util.meta(snr, ds, dop)
xd = x.cuda()
md = tuple(t.cuda() if torch.is_tensor(t) else t for t in md)
with torch.no_grad():
    y = m(xd, md); torch.cuda.synchronize()
    from adafortitran_b200 import _capi
    import ctypes as C
    _capi.lib().aft_profile_enable(m._handle, 1)
    m(xd, md); torch.cuda.synchronize()          # creates the profile events (not part of the timed call)
    st_ms = (C.c_double * 3)(); st_n = (C.c_int64 * 3)()
    _capi.lib().aft_profile_read(m._handle, st_ms, st_n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = m(xd, md); e1.record(); torch.cuda.synchronize()
    st_ms = (C.c_double * 3)(); st_n = (C.c_int64 * 3)()
    _capi.lib().aft_profile_read(m._handle, st_ms, st_n)
ms = e0.elapsed_time(e1)
sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
cfg = O.OracleConfig(num_scs=3276, num_symbols=14, pilot_scs=1638, pilot_symbols=2, patch=(3, 2), num_layers=6, activation="gelu", adaptive=True)
t0 = time.time()
ref = O.forward(cfg, sd, x[:1].numpy(), snr[:1], ds[:1], dop[:1], dtype=np.float64)
t_cpu = time.time() - t0
peak = 1387.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass
enc_tflops = B * 385.5e9 * 0.9967 / (st_ms[1] / 1e3) / 1e12 if st_ms[1] > 0 else None
print(json.dumps({"config": f"AdaFortiTran 3276x14, pilots 1638x2, patch 3x2, S=7644, {m.precision} generic path", "batch": B,
                  "stages_ms": {"frontend": st_ms[0], "encoder": st_ms[1], "head": st_ms[2]},
                  "encoder_tflops": enc_tflops, "encoder_frac_of_sustained_bf16_peak": (enc_tflops / peak) if enc_tflops else None,
                  "rel_err_db_vs_fp64_oracle": O.rel_err_db(y[:1].cpu().numpy(), ref),
                  "ms_per_forward": ms, "estimates_per_s": B / ms * 1e3, "flop_per_estimate": 385.5e9,
                  "tflops_whole_forward": B * 385.5e9 / ms / 1e9, "normwise_err_vs_fp64_oracle": O.normwise_err(y[:1].cpu().numpy(), ref),
                  "oracle_cpu_seconds_per_estimate": t_cpu, "params": sum(p.numel() for p in m.parameters())}))
