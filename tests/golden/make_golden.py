"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs ``/root/reference``; the GPU box does not have it):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference (BerkIGuler/AdaFortiTran) has no tests or golden vectors of its own, so parity is
pinned by executing its ``src.models`` estimators here, on CPU fp32 (its default device), and
recording inputs, weights and outputs.  Files written:

  weights_ada_seed0.npz   state_dict of ``AdaFortiTranEstimator`` built under ``torch.manual_seed(0)``
  golden_ada.npz          B=8 forward (fp32 and .double()), + per-stage tensors for the first 2 samples
  golden_forti.npz        ``FortiTranEstimator`` forward with the same weights (adapter dropped,
                          ``linear_1.weight[:, :6]``)
  golden_variants.npz     sinusoidal pos-enc, relu activation, num_layers=2 outputs (same weights)
  golden_sweep.npz        21-condition synthetic sweep, 4 samples per condition
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AFT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.dont_write_bytecode = True

from src.config.schemas import ModelConfig, SystemConfig  # noqa: E402  (the reference's)
from src.models import AdaFortiTranEstimator, FortiTranEstimator  # noqa: E402

from oracle import aft_oracle as O  # noqa: E402

SYS = dict(ofdm=dict(num_scs=120, num_symbols=14), pilot=dict(num_scs=12, num_symbols=2))
ADA = dict(model_type="adafortitran", patch_size=(3, 2), num_layers=6, model_dim=128, num_head=4,
           activation="gelu", dropout=0.1, max_seq_len=512, pos_encoding_type="learnable",
           channel_adaptivity_hidden_sizes=[7, 42, 560], adaptive_token_length=6)
FORTI = {k: v for k, v in ADA.items() if k not in ("channel_adaptivity_hidden_sizes", "adaptive_token_length")}
FORTI["model_type"] = "fortitran"


def meta(snr, ds, dop):
    b = len(snr)
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32)).reshape(b, 1)
    return (torch.zeros(b, 1), t(snr), t(ds), t(dop), torch.zeros(b, 1), [("syn",) * b])


def forti_state(sd):
    out = {k: v for k, v in sd.items() if not k.startswith("channel_adapter.")}
    out["transformer_encoder.linear_1.weight"] = sd["transformer_encoder.linear_1.weight"][:, :6].clone()
    return out


def main():
    torch.set_num_threads(8)
    sc = SystemConfig(**SYS)
    torch.manual_seed(0)
    ada = AdaFortiTranEstimator(sc, ModelConfig(**ADA)).eval()
    sd = ada.state_dict()
    np.savez(os.path.join(HERE, "weights_ada_seed0.npz"), **{k: v.numpy() for k, v in sd.items()})

    pilots, snr, ds, dop = O.synthetic_batch(8, seed=1)
    tp = torch.from_numpy(pilots)
    with torch.no_grad():
        out32 = ada(tp, meta(snr, ds, dop)).numpy()
        # stage tensors through the reference's own sub-modules, first two samples, real part
        x = tp.real[:2]
        up = ada.pilot_upsampler(x.reshape(2, -1))
        enh = ada.initial_enhancer(up.view(2, 1, 120, 14)).squeeze(1)
        tok = ada.patch_embedder(enh)
        cond = ada.channel_adapter(*[t[:2] for t in meta(snr, ds, dop)[1:4]])
        tok_in = torch.cat((tok, cond), dim=2)
        enc = ada.transformer_encoder
        h0 = enc.positional_encoding(enc.linear_1(tok_in))
        hs = [h0]
        for layer in enc.transformer.layers:
            hs.append(layer(hs[-1]))
        tok_out = enc.linear_2(hs[-1])
        comb = enh + ada.patch_reconstructor(tok_out)
        ada64 = AdaFortiTranEstimator(sc, ModelConfig(**ADA)).double().eval()
        ada64.load_state_dict({k: v.double() for k, v in sd.items()})
        m64 = tuple(t.double() if torch.is_tensor(t) else t for t in meta(snr, ds, dop))
        out64 = ada64(tp.to(torch.complex128), m64).numpy()
    np.savez(os.path.join(HERE, "golden_ada.npz"), pilots=pilots, snr=snr, ds=ds, dop=dop, out=out32, out64=out64,
             st_upsampled=up.numpy(), st_conv_enhanced=enh.numpy(), st_tokens=tok_in.numpy(),
             st_adapter=cond.numpy(), st_h0=h0.numpy(), st_h1=hs[1].numpy(), st_h6=hs[6].numpy(),
             st_tok_out=tok_out.numpy(), st_combined=comb.numpy())

    forti = FortiTranEstimator(sc, ModelConfig(**FORTI)).eval()
    forti.load_state_dict(forti_state(sd))
    with torch.no_grad():
        outf = forti(tp).numpy()
    np.savez(os.path.join(HERE, "golden_forti.npz"), pilots=pilots, out=outf)

    var = {}
    with torch.no_grad():
        m = AdaFortiTranEstimator(sc, ModelConfig(**{**ADA, "pos_encoding_type": "sinusoidal"})).eval()
        s2 = {k: v for k, v in sd.items() if "position_embeddings" not in k}
        s2["transformer_encoder.positional_encoding.pe"] = m.state_dict()["transformer_encoder.positional_encoding.pe"]
        m.load_state_dict(s2)
        var["out_sinusoidal"] = m(tp[:4], meta(snr[:4], ds[:4], dop[:4])).numpy()
        var["pe_first_rows"] = s2["transformer_encoder.positional_encoding.pe"][0, :280].numpy()
        m = AdaFortiTranEstimator(sc, ModelConfig(**{**ADA, "activation": "relu"})).eval()
        m.load_state_dict(sd)
        var["out_relu"] = m(tp[:4], meta(snr[:4], ds[:4], dop[:4])).numpy()
        m = AdaFortiTranEstimator(sc, ModelConfig(**{**ADA, "num_layers": 2})).eval()
        m.load_state_dict({k: v for k, v in sd.items() if not any(f"layers.{i}." in k for i in range(2, 6))})
        var["out_layers2"] = m(tp[:4], meta(snr[:4], ds[:4], dop[:4])).numpy()
    np.savez(os.path.join(HERE, "golden_variants.npz"), **var)

    # 21-condition sweep: SNR / DS / Doppler swept one at a time around 20 / 50 / 500
    conds = [(s, 50.0, 500.0) for s in O.SNR_GRID] + [(20.0, d, 500.0) for d in O.DS_GRID] + \
            [(20.0, 50.0, f) for f in O.DOP_GRID]
    outs, ps, hs_, cs = [], [], [], []
    with torch.no_grad():
        for i, (s, d, f) in enumerate(conds):
            p, h = O.synthetic_channel(4, float(s), float(d), float(f), seed=4242 + i)
            o = ada(torch.from_numpy(p), meta([s] * 4, [d] * 4, [f] * 4)).numpy()
            outs.append(o), ps.append(p), hs_.append(h), cs.append((s, d, f))
    np.savez(os.path.join(HERE, "golden_sweep.npz"), conds=np.asarray(cs, dtype=np.float32),
             pilots=np.stack(ps), truth=np.stack(hs_).astype(np.complex64), out=np.stack(outs))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
