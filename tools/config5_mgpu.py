"""BASELINE configs[4] (AdaFortiTran at 3276 x 14, S = 7644, bf16 long-sequence path) on N GPUs of one node: every rank runs
its own batch (independent replicas -- the path has no cross-sample arithmetic), barrier + CUDA events, max over ranks.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/config5_mgpu.py [B_per_gpu] [steps]
Prints one JSON line on rank 0 (whole-job estimates/s, fraction of the measured sustained bf16 peak)."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafortitran_b200 import AdaFortiTranEstimator, ModelConfig, SystemConfig
from tests import util

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sysc = SystemConfig(ofdm=dict(num_scs=3276, num_symbols=14), pilot=dict(num_scs=1638, num_symbols=2))
modc = ModelConfig(model_type="adafortitran", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4, activation="gelu",
                   max_seq_len=7644, pos_encoding_type="learnable", channel_adaptivity_hidden_sizes=[7, 42, 15288],
                   adaptive_token_length=6, device=f"cuda:{local}")
torch.manual_seed(0)
m = AdaFortiTranEstimator(sysc, modc).eval()
m.precision = "bf16"
g = torch.Generator().manual_seed(3 + rank)
x = torch.complex(torch.randn(B, 1638, 2, generator=g), torch.randn(B, 1638, 2, generator=g)).to(dev)
md = tuple(t.to(dev) if torch.is_tensor(t) else t for t in util.meta(np.full(B, 20.0, np.float32), np.full(B, 50.0, np.float32), np.full(B, 500.0, np.float32)))
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
with torch.no_grad():
    for _ in range(2):
        y = m(x, md)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y = m(x, md)
    e1.record()
    barrier()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
finite = bool(torch.isfinite(torch.view_as_real(y)).all())
if rank == 0:
    F_EST = 385_467_110_364          # FLOP per estimate at this geometry
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peaks = json.load(open(pk)) if os.path.exists(pk) else {}
    sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
    value = world * B * steps / (float(ms.item()) / 1e3)
    print(json.dumps({"config": "BASELINE configs[4]: AdaFortiTran 3276x14 (S=7644), bf16 long-sequence path, independent replicas",
                      "n_gpus": world, "batch_per_gpu": B, "steps": steps, "ms_per_step": float(ms.item()) / steps,
                      "estimates_per_s": value, "tflops_per_gpu": value / world * F_EST / 1e12,
                      "frac_of_sustained_bf16_peak": value / world * F_EST / 1e12 / sustained, "peak_tflops": sustained, "finite": finite}), flush=True)
if world > 1:
    dist.destroy_process_group()
