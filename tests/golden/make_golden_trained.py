"""A "pseudo-trained" weight set for the bf16 accuracy gates.

At random initialisation every output of the network gives NMSE ~ 0 dB against a unit-power channel, so the north-star
gate |dNMSE| <= 0.05 dB cannot fail.  This script trains the LIVE reference (``/root/reference``: ``AdaFortiTranEstimator``,
its own forward / parameters, CPU fp32) for a few hundred Adam steps on the synthetic doubly-selective channels of
``oracle.aft_oracle.synthetic_channel`` with the reference's loss (MSE over concatenated real / imaginary parts,
``src/main/trainer.py:161-170`` with ``src/utils.py:164-180``), until the NMSE is well below 0 dB, and records

  golden_trained.npz : the trained ``state_dict`` (fp32), the reference's fp32 outputs on the 21-condition sweep (4 samples
                       per condition; pilots and channel truth are NOT stored: the first 4 samples of
                       ``O.synthetic_channel(8, snr, ds, dop, seed=9000 + i)`` regenerate them bit for bit), its NMSE per condition and the training NMSE trace.

Run in the build container only (needs /root/reference):  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_trained.py
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AFT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.dont_write_bytecode = True

from src.config.schemas import ModelConfig, SystemConfig  # noqa: E402  (the reference's)
from src.models import AdaFortiTranEstimator  # noqa: E402

from oracle import aft_oracle as O  # noqa: E402

SYS = dict(ofdm=dict(num_scs=120, num_symbols=14), pilot=dict(num_scs=12, num_symbols=2))
ADA = dict(model_type="adafortitran", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4,
           activation="gelu", dropout=0.1, max_seq_len=512, pos_encoding_type="learnable",
           channel_adaptivity_hidden_sizes=[7, 42, 560], adaptive_token_length=6)
STEPS = int(os.environ.get("AFT_TRAIN_STEPS", "400"))
BATCH = 32


def meta(snr, ds, dop):
    b = len(snr)
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32)).reshape(b, 1)
    return (torch.zeros(b, 1), t(snr), t(ds), t(dop), torch.zeros(b, 1), [("syn",) * b])


def cat_ri(x):   # the reference's concat_complex_channel (src/utils.py:164-180)
    return torch.cat([x.real, x.imag], dim=1)


def batch(rng, n, seed):
    """n samples, each with its own condition from the reference's 7 x 7 x 7 grid."""
    snr = rng.choice(O.SNR_GRID, size=n).astype(np.float32)
    ds = rng.choice(O.DS_GRID, size=n).astype(np.float32)
    dop = rng.choice(O.DOP_GRID, size=n).astype(np.float32)
    ps, hs = [], []
    for i in range(n):
        p, h = O.synthetic_channel(1, float(snr[i]), float(ds[i]), float(dop[i]), seed=seed * 1000 + i)
        ps.append(p[0]), hs.append(h[0])
    return np.stack(ps), np.stack(hs), snr, ds, dop


def main():
    torch.set_num_threads(8)
    torch.manual_seed(0)
    model = AdaFortiTranEstimator(SystemConfig(**SYS), ModelConfig(**ADA))
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    loss_fn = torch.nn.MSELoss()
    rng = np.random.default_rng(7)
    trace = []
    t0 = time.time()
    model.train()
    for step in range(STEPS):
        p, h, snr, ds, dop = batch(rng, BATCH, step)
        est = model(torch.from_numpy(p), meta(snr, ds, dop))
        truth = torch.from_numpy(h)
        loss = loss_fn(cat_ri(est), cat_ri(truth))
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        nmse = 10 * np.log10(float((est - truth).abs().pow(2).sum() / truth.abs().pow(2).sum()))
        trace.append(nmse)
        if step % 25 == 0 or step == STEPS - 1:
            print(f"step {step:4d} loss {float(loss):.5f} nmse {nmse:7.2f} dB  ({time.time() - t0:.0f} s)", flush=True)
    model.eval()
    conds = [(s, 50.0, 500.0) for s in O.SNR_GRID] + [(20.0, d, 500.0) for d in O.DS_GRID] + [(20.0, 50.0, f) for f in O.DOP_GRID]
    outs, ps, hs, cs = [], [], [], []
    with torch.no_grad():
        for i, (s, d, f) in enumerate(conds):
            p, h = O.synthetic_channel(8, float(s), float(d), float(f), seed=9000 + i)
            p, h = p[:4], h[:4]
            o = model(torch.from_numpy(p), meta([s] * 4, [d] * 4, [f] * 4)).numpy()
            outs.append(o), ps.append(p), hs.append(h), cs.append((s, d, f))
    out, truth = np.stack(outs), np.stack(hs).astype(np.complex64)
    nm = [O.nmse_db(out[i], truth[i]) for i in range(len(conds))]
    print("eval NMSE per condition (dB):", " ".join(f"{v:.2f}" for v in nm))
    sd = {"sd/" + k: v.detach().numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "golden_trained.npz"), conds=np.asarray(cs, dtype=np.float32), out=out,
                        nmse_db=np.asarray(nm, dtype=np.float32), train_nmse_db=np.asarray(trace, dtype=np.float32),
                        samples=np.int64(4), gen_batch=np.int64(8), seed0=np.int64(9000), **sd)
    print("written", os.path.join(HERE, "golden_trained.npz"))


if __name__ == "__main__":
    main()
