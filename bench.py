#!/usr/bin/env python
"""bench.py -- channel estimates / second of the AdaFortiTran/FortiTran inference forward on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU arithmetic on the host cores

One "step" = one forward pass over one synthetic batch.  Workload at every N (weak scaling, one independent
replica per GPU, no data-path collective): BASELINE.json configs[1] -- FortiTran default config, bf16, batch
16384 per GPU, random-init weights, unit-power complex Gaussian pilots.  `--workload ada` switches to the
AdaFortiTran default config (configs[2] shape) with SNR/DS/Doppler metadata.

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM (CUDA events on the launch
stream, one event pair per step, L2 flushed between steps, max over ranks); `e2e` goes through the
host-buffer C-ABI entry point (pinned host memory in, pinned host memory out, copies inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_EST = {"forti": 1_428_241_920, "ada": 1_429_387_932}   # algorithmic FLOPs / complex estimate
F_ENC_EST = 1_362_493_440                                 # of which encoder GEMM + attention (the encoder kernel)
BYTES_IN = {"forti": 192, "ada": 204}
BYTES_OUT = 13_440
SYS = dict(ofdm=dict(num_scs=120, num_symbols=14), pilot=dict(num_scs=12, num_symbols=2))
FORTI = dict(model_type="fortitran", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4, activation="gelu",
             dropout=0.1, max_seq_len=512, pos_encoding_type="learnable")
ADA = dict(FORTI, model_type="adafortitran", channel_adaptivity_hidden_sizes=[7, 42, 560], adaptive_token_length=6)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--workload", choices=("forti", "ada"), default="forti")
    ap.add_argument("--batch-per-gpu", type=int, default=16384)
    ap.add_argument("--precision", choices=("bf16", "fp32"), default="bf16")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), hbm=p["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_rate(kind, seconds, warmup=1, steps=None, batch=64):
    """Reference arithmetic on the host cores (oracle/torch_port.py: the reference's composition over the same
    torch operators, bit-identical to the reference on the golden vectors).  Bounded sample."""
    import numpy as np
    import torch
    from oracle import aft_oracle as O
    from oracle.torch_port import TorchPort
    from adafortitran_b200 import AdaFortiTranEstimator, FortiTranEstimator, ModelConfig, SystemConfig
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    cls, cfg = (AdaFortiTranEstimator, ADA) if kind == "ada" else (FortiTranEstimator, FORTI)
    sd = {k: v.numpy() for k, v in cls(SystemConfig(**SYS), ModelConfig(**cfg)).state_dict().items()}
    port = TorchPort(sd, adaptive=(kind == "ada"))
    p, snr, ds, dop = O.synthetic_batch(batch, seed=1)
    args = (torch.from_numpy(p),) + ((snr, ds, dop) if kind == "ada" else ())
    for _ in range(warmup):
        port(*args)
    times = []
    t_end = time.perf_counter() + seconds
    while (steps is None and time.perf_counter() < t_end and len(times) < 200) or (steps is not None and len(times) < steps):
        t0 = time.perf_counter()
        port(*args)
        times.append(time.perf_counter() - t0)
        if steps is None and len(times) >= 3 and time.perf_counter() >= t_end:
            break
    total = sum(times)
    return {"value": batch * len(times) / total, "unit": "estimates/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} forward passes of batch {batch} ({kind} default config, fp32, torch {torch.__version__} "
                      f"CPU, {torch.get_num_threads()} threads), {total:.1f} s",
            "ms_per_step": 1e3 * total / len(times), "steps": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_rate(args.workload, args.cpu_seconds, warmup=max(1, args.warmup), steps=args.steps)
    line = {
        "impl": "reference", "metric": "channel_estimates_per_sec", "value": r["value"], "unit": "estimates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, note="reference arm: each step is a bounded sample of 64 estimates on the host CPU"),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "estimates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, note=None):
    name = "FortiTran" if args.workload == "forti" else "AdaFortiTran"
    cfg = {"workload": f"{name} default config (120x14 grid, 12x2 pilots, patch 3x2, 6 layers, d=128, 4 heads), "
                       f"{args.precision} inference, batch {args.batch_per_gpu} per GPU, random-init weights, synthetic "
                       f"CN(0,1) pilots" + (" + SNR/DS/Doppler metadata" if args.workload == "ada" else ""),
           "batch_per_gpu": args.batch_per_gpu, "global_batch": args.batch_per_gpu * args.gpus,
           "parallelism": f"{args.gpus} independent replica(s), batch-sharded, no data-path collective",
           "l2": "L2 flushed between timed steps (256 MiB device write); per-step intermediate stream (2.4 GB) exceeds the 126 MB L2"}
    if note:
        cfg["note"] = note
    return cfg


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ctypes as C
    from adafortitran_b200 import AdaFortiTranEstimator, FortiTranEstimator, ModelConfig, SystemConfig, _capi
    from adafortitran_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch_per_gpu
    kind = args.workload
    torch.manual_seed(0)   # identical random-init weights on every replica
    cls, cfg = (AdaFortiTranEstimator, ADA) if kind == "ada" else (FortiTranEstimator, FORTI)
    model = cls(SystemConfig(**SYS), ModelConfig(**dict(cfg, device=f"cuda:{local}"))).eval()
    model.precision = args.precision

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    pilots_h = torch.view_as_complex(torch.randn(B, 12, 2, 2, generator=g) * (0.5 ** 0.5)).contiguous().pin_memory()
    meta_h = None
    if kind == "ada":
        pick = lambda grid: torch.tensor(grid, dtype=torch.float32)[torch.randint(0, 7, (B,), generator=g)].reshape(B, 1).pin_memory()
        meta_h = (torch.zeros(B, 1), pick(list(range(0, 31, 5))), pick(list(range(50, 351, 50))), pick(list(range(200, 1401, 200))),
                  torch.zeros(B, 1), None)
    pilots_d = pilots_h.to(dev)
    meta_d = None if meta_h is None else tuple(t.to(dev) if torch.is_tensor(t) else t for t in meta_h)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            out = model(pilots_d, meta_d)
        barrier()
        lib = _capi.lib()
        lib.aft_profile_enable(model._handle, 1)
        launches0 = lib.aft_launch_count()
        sampler = ClockSampler(local)
        sampler.start()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        barrier()
        wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()                      # evict L2 (untimed)
            starts[i].record()
            out = model(pilots_d, meta_d)
            stops[i].record()
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.summary()
        step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
        launches = lib.aft_launch_count() - launches0
        ms = (C.c_double * 3)()
        nl = (C.c_int64 * 3)()
        _capi.check(lib.aft_profile_read(model._handle, ms, nl))
        lib.aft_profile_enable(model._handle, 0)
        total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        total_ms = float(total_ms.item())
        value = world * B * args.steps / (total_ms / 1e3)

        # ---- end to end through the host-buffer entry point (pinned in, pinned out) ----
        out_h = torch.empty((B, 120, 14), dtype=torch.complex64).pin_memory()
        model.forward_host(pilots_h, meta_h, out=out_h)    # warm-up (allocates the internal lanes)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            model.forward_host(pilots_h, meta_h, out=out_h)
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_value = world * B * args.e2e_steps / float(e2e_s.item())
        same = bool(torch.equal(out_h.to(dev), out))

        # ---- the only collectives of the path: all-gather of the estimates + reduction of the error sums ----
        coll = None
        sums = torch.zeros(2, dtype=torch.float64, device=dev)
        truth = torch.zeros_like(out)
        st = torch.cuda.current_stream().cuda_stream
        _capi.check(lib.aft_error_sums(C.c_void_p(out.data_ptr()), C.c_void_p(truth.data_ptr()), out.numel(),
                                      C.c_void_p(sums.data_ptr()), C.c_void_p(st)))
        if world > 1:
            warm = D.gather_estimates(out)         # NCCL warm-up at full size (communicator / buffer setup is not part of the path)
            del warm
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gathered = D.gather_estimates(out)
            D.reduce_error_sums(sums)
            e1.record()
            torch.cuda.synchronize()
            coll = {"all_gather_plus_all_reduce_ms": e0.elapsed_time(e1), "all_gather_bytes_per_rank": out.numel() * 8,
                    "gathered_shape": list(gathered.shape)}
            del gathered
        out_power = float(sums[0].item()) / (world * out.numel())

    pk = peaks()
    enc_ms = ms[1]
    n_est = B * args.steps
    roof = None
    if enc_ms > 0 and args.precision == "bf16":
        ach = n_est * F_ENC_EST / (enc_ms / 1e3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "encoder_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        roof = {"bound": "tensor", "kernel": "encoder_kernel", "achieved": ach, "peak": pk["sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["sustained"], "traffic": traffic, "peak_source": pk["source"] + ", sustained figure",
                "launches": int(nl[1]), "avg_launch_ms": enc_ms / max(1, int(nl[1])),
                "algorithmic_flops_per_launch": n_est * F_ENC_EST / max(1, int(nl[1]))}
    line = {
        "metric": "channel_estimates_per_sec", "value": value, "unit": "estimates/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "estimates/s", "h2d_bytes_per_step": B * BYTES_IN[kind],
                "d2h_bytes_per_step": B * BYTES_OUT, "steps": args.e2e_steps, "matches_device_path": same},
        "gpu_launches": int(launches),
        "roofline": roof,
        "step_tensor_frac": value / world * F_EST[kind] / 1e12 / pk["sustained"],
        "stages_ms_per_step": {"frontend": ms[0] / args.steps, "encoder": ms[1] / args.steps, "head": ms[2] / args.steps},
        "wall_s_timed_region": wall, "output_mean_power": out_power, "collectives": coll,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_rate(kind, args.cpu_seconds)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
