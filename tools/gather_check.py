"""Multi-rank check of the fused all-gather (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/gather_check.py
Every rank forwards its shard through ShardedEvaluator in "peer" mode (head kernel stores into all gather buffers) and in
"nccl" mode (forward + all_gather_into_tensor); the gathered arrays must be bit-identical, equal on every rank, and
equal to a plain forward of the whole batch.  Also times both modes.  Prints one JSON line on rank 0."""
import json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafortitran_b200 import distributed as D
from oracle import aft_oracle as O          # synthetic inputs only
from tests import util


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    B = per * world
    sd = util.ada_weights()
    p, snr, ds, dop = O.synthetic_batch(B, seed=11)
    model = util.make_model("ada", device=f"cuda:{local}", weights=sd, precision=precision)
    lo, hi = D.shard_range(B, rank, world)
    pil = torch.from_numpy(p[lo:hi]).to(dev)
    md = tuple(t.to(dev) if torch.is_tensor(t) else t for t in util.meta(snr[lo:hi], ds[lo:hi], dop[lo:hi]))
    truth = torch.zeros((hi - lo, 120, 14), dtype=torch.complex64, device=dev)
    res = {"world": world, "per_rank": per, "precision": precision}
    with torch.no_grad():
        full = model(torch.from_numpy(p).to(dev), tuple(t.to(dev) if torch.is_tensor(t) else t for t in util.meta(snr, ds, dop)))
        out = {}
        for mode in ("peer", "nccl"):
            ev = D.ShardedEvaluator(model, per, mode=mode)
            for _ in range(2):
                g, s = ev.step(pil, md, truth)
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                g, s = ev.step(pil, md, truth)
            torch.cuda.synchronize(); dist.barrier()
            res[f"{mode}_ms"] = (time.perf_counter() - t0) / 5 * 1e3
            out[mode] = (g.clone(), s.clone())
            # host path: pinned inputs in, local estimates out
            ph = torch.from_numpy(p[lo:hi]).pin_memory()
            mh = util.meta(snr[lo:hi], ds[lo:hi], dop[lo:hi])
            oh = torch.empty((hi - lo, 120, 14), dtype=torch.complex64).pin_memory()
            g2, s2 = ev.step(ph, mh, truth, host_out=oh)
            torch.cuda.synchronize(); dist.barrier()
            res[f"{mode}_host_equal"] = bool(torch.equal(g2, out[mode][0])) and bool(torch.equal(oh.to(dev), full[lo:hi]))
            ev.close()
        res["peer_equals_nccl"] = bool(torch.equal(out["peer"][0], out["nccl"][0]))
        res["gathered_equals_full_forward"] = bool(torch.equal(out["peer"][0], full))
        res["sums_equal"] = bool(torch.allclose(out["peer"][1], out["nccl"][1], rtol=1e-12))
        ok = torch.tensor([int(res["peer_equals_nccl"] and res["gathered_equals_full_forward"] and res["peer_host_equal"])], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res["all_ranks_ok"] = bool(ok.item())
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
