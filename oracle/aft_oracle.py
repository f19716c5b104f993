"""CPU oracle for the AdaFortiTran / FortiTran inference forward pass.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / reported CPU
baseline.  The product path (``adafortitran_b200``) never imports this module and
fails loudly when its CUDA extension is missing.

Parity pinning: the reference ships no tests or golden vectors of its own
(SURVEY.md §4, §8c).  This restatement is pinned against outputs of the live
reference, executed in the build container from ``/root/reference`` by
``tests/golden/make_golden.py`` (committed), whose inputs/weights/outputs are stored
under ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this file
against those vectors (end-to-end and stage by stage).  The same holds for the rows
either side of the path: ``make_golden_next.py`` records the reference's ``MatDataset`` /
``extract_values`` / ``LinearEstimator`` (``tests/test_next_rows.py``) and
``make_golden_generic.py`` its forward at two non-default grids
(``tests/test_generic_grid.py``).

What it restates (all citations relative to the reference tree):
  * ``src/models/fortitran.py:145-182``   complex -> two real passes -> complex
  * ``src/models/fortitran.py:184-233``   the 8-stage real-valued pipeline
  * ``src/models/blocks/enhancers.py:12-20``          ConvEnhancer 1->8->32->8->1, 3x3, pad 1
  * ``src/models/blocks/patch_processors.py:34-35``   Unfold(k=s=patch) + permute
  * ``src/models/blocks/patch_processors.py:69-71``   permute + Fold (exact inverse)
  * ``src/models/blocks/channel_adaptivity.py:34-40,59-63``  three 1->h1->h2->h3 MLPs
  * ``src/models/blocks/encoders.py:67-70``           linear_1, pos-enc, encoder, linear_2
  * ``src/models/blocks/positional_encodings.py:38,64`` additive table, first S rows
  * torch ``nn.TransformerEncoderLayer`` defaults used at ``encoders.py:44-51``:
    post-norm, LayerNorm eps 1e-5 (biased variance), exact-erf GELU (or ReLU),
    dropout inactive in eval, softmax over keys of q.k/sqrt(head_dim).

Arithmetic is plain numpy in the dtype requested (float64 by default, float32 to
mimic the reference's working precision).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

try:  # scipy is present in the image; keep a pure-python fallback so the oracle never disappears
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])


@dataclass(frozen=True)
class OracleConfig:
    """Shape parameters of one model instance (``config/*.yaml`` + ``system_config.yaml``)."""

    num_scs: int = 120
    num_symbols: int = 14
    pilot_scs: int = 12
    pilot_symbols: int = 2
    patch: Tuple[int, int] = (3, 2)
    num_layers: int = 6
    model_dim: int = 128
    num_head: int = 4
    activation: str = "gelu"
    adaptive: bool = True

    @property
    def tokens(self) -> int:
        return (self.num_scs // self.patch[0]) * (self.num_symbols // self.patch[1])

    @property
    def patch_len(self) -> int:
        return self.patch[0] * self.patch[1]


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------

def linear(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray]) -> np.ndarray:
    """``y = x @ w.T + b`` (torch ``nn.Linear`` semantics; ``w`` is ``[out, in]``)."""
    y = x @ w.T
    return y if b is None else y + b


def conv3x3(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """3x3 cross-correlation, stride 1, zero padding 1 (``nn.Conv2d(..., 3, padding=1)``).

    ``x``: [N, Cin, H, W]; ``w``: [Cout, Cin, 3, 3]; returns [N, Cout, H, W].
    Reference: ``enhancers.py:13-19``.
    """
    n, cin, h, wd = x.shape
    xp = np.zeros((n, cin, h + 2, wd + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((n, w.shape[0], h, wd), dtype=x.dtype)
    for dy in range(3):
        for dx in range(3):
            # [N,Cin,H,W] x [Cout,Cin] -> [N,Cout,H,W]
            out += np.einsum("nchw,oc->nohw", xp[:, :, dy:dy + h, dx:dx + wd], w[:, :, dy, dx])
    return out + b[None, :, None, None]


def conv_enhancer(x: np.ndarray, sd: Dict[str, np.ndarray], prefix: str) -> np.ndarray:
    """ConvEnhancer: conv-ReLU-conv-ReLU-conv-ReLU-conv (``enhancers.py:12-20``)."""
    y = x
    for i, idx in enumerate((0, 2, 4, 6)):
        y = conv3x3(y, sd[f"{prefix}.conv_block.{idx}.weight"], sd[f"{prefix}.conv_block.{idx}.bias"])
        if i < 3:
            y = np.maximum(y, 0)
    return y


def patchify(img: np.ndarray, patch: Tuple[int, int]) -> np.ndarray:
    """[N,H,W] -> [N,T,ph*pw]; token t = i*(W/pw)+j, feature f = a*pw+b <-> pixel (ph*i+a, pw*j+b).

    Equivalent of ``Unfold(kernel=stride=patch)`` + ``permute(0,2,1)`` (``patch_processors.py:34-35``).
    """
    n, h, w = img.shape
    ph, pw = patch
    x = img.reshape(n, h // ph, ph, w // pw, pw)       # [N,i,a,j,b]
    x = x.transpose(0, 1, 3, 2, 4)                      # [N,i,j,a,b]
    return x.reshape(n, (h // ph) * (w // pw), ph * pw)


def unpatchify(tok: np.ndarray, grid: Tuple[int, int], patch: Tuple[int, int]) -> np.ndarray:
    """Exact inverse of :func:`patchify` (``Fold`` with non-overlapping patches, ``patch_processors.py:69-71``)."""
    n = tok.shape[0]
    h, w = grid
    ph, pw = patch
    x = tok.reshape(n, h // ph, w // pw, ph, pw).transpose(0, 1, 3, 2, 4)
    return x.reshape(n, h, w)


def adapter(sd: Dict[str, np.ndarray], snr: np.ndarray, ds: np.ndarray, dop: np.ndarray, tokens: int) -> np.ndarray:
    """ChannelAdapter (``channel_adaptivity.py:34-40,59-63``): [B,1]x3 -> [B,T,6].

    Feature order per token: [snr0, snr1, ds0, ds1, dop0, dop1]; token t takes MLP outputs [2t, 2t+1].
    """
    outs = []
    for name, v in (("snr", snr), ("ds", ds), ("dop", dop)):
        p = f"channel_adapter.{name}_encoder"
        z = v.reshape(-1, 1)
        z = np.maximum(linear(z, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"]), 0)
        z = np.maximum(linear(z, sd[f"{p}.2.weight"], sd[f"{p}.2.bias"]), 0)
        z = linear(z, sd[f"{p}.4.weight"], sd[f"{p}.4.bias"])
        outs.append(z.reshape(z.shape[0], tokens, -1))
    return np.concatenate(outs, axis=2)


def layer_norm(x: np.ndarray, g: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)   # biased, as torch
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * g + b


def gelu_erf(x: np.ndarray) -> np.ndarray:
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def encoder_layer(h: np.ndarray, sd: Dict[str, np.ndarray], l: int, num_head: int, activation: str) -> np.ndarray:
    """One post-norm ``nn.TransformerEncoderLayer`` in eval mode (``encoders.py:44-51``)."""
    p = f"transformer_encoder.transformer.layers.{l}"
    n, s, d = h.shape
    dh = d // num_head
    qkv = linear(h, sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"])
    q, k, v = (qkv[..., i * d:(i + 1) * d].reshape(n, s, num_head, dh).transpose(0, 2, 1, 3) for i in range(3))
    sc = (q @ k.transpose(0, 1, 3, 2)) * h.dtype.type(1.0 / math.sqrt(dh))
    sc = sc - sc.max(axis=-1, keepdims=True)
    e = np.exp(sc)
    a = e / e.sum(axis=-1, keepdims=True)
    o = (a @ v).transpose(0, 2, 1, 3).reshape(n, s, d)
    o = linear(o, sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"])
    h = layer_norm(h + o, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"])
    f = linear(h, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])
    f = gelu_erf(f) if activation == "gelu" else np.maximum(f, 0)
    f = linear(f, sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
    return layer_norm(h + f, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"])


def pos_table(sd: Dict[str, np.ndarray]) -> np.ndarray:
    """The additive positional table, whichever variant the state_dict carries."""
    for key in ("transformer_encoder.positional_encoding.position_embeddings",
                "transformer_encoder.positional_encoding.pe"):
        if key in sd:
            return sd[key][0]
    raise KeyError("no positional table in state_dict")


# --------------------------------------------------------------------------------------
# the pipeline
# --------------------------------------------------------------------------------------

def forward_real(cfg: OracleConfig, sd: Dict[str, np.ndarray], x: np.ndarray,
                 cond: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None,
                 stages: Optional[dict] = None) -> np.ndarray:
    """One real-valued pass, ``fortitran.py:184-233``.  ``x``: [N, pilot_scs, pilot_symbols]."""
    n = x.shape[0]
    grid = (cfg.num_scs, cfg.num_symbols)
    up = linear(x.reshape(n, -1), sd["pilot_upsampler.weight"], sd["pilot_upsampler.bias"])
    img = up.reshape(n, 1, *grid)
    enh = conv_enhancer(img, sd, "initial_enhancer")[:, 0]
    tok = patchify(enh, cfg.patch)
    if cfg.adaptive:
        assert cond is not None
        ada = adapter(sd, *cond, cfg.tokens).astype(x.dtype)
        tok_in = np.concatenate([tok, ada], axis=2)
    else:
        tok_in = tok
    h = linear(tok_in, sd["transformer_encoder.linear_1.weight"], sd["transformer_encoder.linear_1.bias"])
    h = h + pos_table(sd)[: h.shape[1]]
    if stages is not None:
        stages.update(upsampled=up, conv_enhanced=enh, tokens=tok_in, h0=h)
    for l in range(cfg.num_layers):
        h = encoder_layer(h, sd, l, cfg.num_head, cfg.activation)
        if stages is not None:
            stages[f"h{l + 1}"] = h
    r = linear(h, sd["transformer_encoder.linear_2.weight"], sd["transformer_encoder.linear_2.bias"])
    rec = unpatchify(r, grid, cfg.patch)
    comb = enh + rec
    out = conv_enhancer(comb[:, None], sd, "final_refiner")[:, 0]
    if stages is not None:
        stages.update(tok_out=r, combined=comb, refined=out)
    return out


def forward(cfg: OracleConfig, sd: Dict[str, np.ndarray], pilots: np.ndarray,
            snr: Optional[np.ndarray] = None, ds: Optional[np.ndarray] = None, dop: Optional[np.ndarray] = None,
            dtype=np.float64, stages: Optional[dict] = None) -> np.ndarray:
    """Complex forward, ``fortitran.py:145-182``: same real network on ``.real`` and ``.imag``.

    ``pilots``: complex [B, pilot_scs, pilot_symbols]; ``snr/ds/dop``: [B] or [B,1] (adaptive only).
    Returns complex [B, num_scs, num_symbols] (complex128 for float64, complex64 for float32).
    """
    sd = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    cond = None
    if cfg.adaptive:
        if snr is None or ds is None or dop is None:
            raise ValueError("meta_data is required when channel adaptation is enabled")  # fortitran.py:157-158
        cond = tuple(np.asarray(v, dtype=dtype).reshape(-1) for v in (snr, ds, dop))
    pilots = np.asarray(pilots)
    st_re = {} if stages is not None else None
    st_im = {} if stages is not None else None
    re = forward_real(cfg, sd, pilots.real.astype(dtype), cond, st_re)
    im = forward_real(cfg, sd, pilots.imag.astype(dtype), cond, st_im)
    if stages is not None:
        stages["re"], stages["im"] = st_re, st_im
    return re + 1j * im if dtype == np.float64 else (re + 1j * im).astype(np.complex64)


# --------------------------------------------------------------------------------------
# metrics (reference definitions)
# --------------------------------------------------------------------------------------

def mse_db_reference(est: np.ndarray, truth: np.ndarray) -> float:
    """The reference's reported test metric: ``to_db(2 * MSELoss(cat(re,im)))`` averaged per sample
    (``trainer.py:338-345``, ``utils.py:164-180,233-245``) == 10 log10(mean |est-truth|^2)."""
    return float(10 * np.log10(np.mean(np.abs(est - truth) ** 2)))


def nmse_db(est: np.ndarray, truth: np.ndarray) -> float:
    return float(10 * np.log10(np.sum(np.abs(est - truth) ** 2) / np.sum(np.abs(truth) ** 2)))


def normwise_err(y: np.ndarray, ref: np.ndarray) -> float:
    """max|y-ref| / max|ref| -- the 1e-4 fp32 parity gate."""
    return float(np.max(np.abs(y - ref)) / np.max(np.abs(ref)))


def rel_err_db(y: np.ndarray, ref: np.ndarray) -> float:
    """Output-relative error power in dB: 10 log10(sum|y-ref|^2 / sum|ref|^2)."""
    return float(10 * np.log10(np.sum(np.abs(y - ref) ** 2) / np.sum(np.abs(ref) ** 2)))


# --------------------------------------------------------------------------------------
# synthetic inputs (shared by tests / bench so every leg sees the same data)
# --------------------------------------------------------------------------------------

SNR_GRID = np.arange(0, 31, 5, dtype=np.float32)          # README.md:184-207
DS_GRID = np.arange(50, 351, 50, dtype=np.float32)
DOP_GRID = np.arange(200, 1401, 200, dtype=np.float32)


def synthetic_batch(batch: int, seed: int = 1, cfg: OracleConfig = OracleConfig()):
    """Unit-power CN(0,1) pilots + metadata drawn from the reference's 7x7x7 condition grid."""
    rng = np.random.default_rng(seed)
    shape = (batch, cfg.pilot_scs, cfg.pilot_symbols)
    pilots = ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / math.sqrt(2)).astype(np.complex64)
    snr = rng.choice(SNR_GRID, size=batch).astype(np.float32)
    ds = rng.choice(DS_GRID, size=batch).astype(np.float32)
    dop = rng.choice(DOP_GRID, size=batch).astype(np.float32)
    return pilots, snr, ds, dop


def synthetic_channel(batch: int, snr_db: float, ds_ns: float, dop_hz: float, seed: int = 4242,
                      cfg: OracleConfig = OracleConfig(), taps: int = 8):
    """Synthetic doubly-selective channel + LS pilots (our generator, the
    reference ships none).  Returns (pilots c64 [B,12,2], truth c64 [B,120,14])."""
    rng = np.random.default_rng(seed)
    k = np.arange(cfg.num_scs)[None, None, :, None]
    l = np.arange(cfg.num_symbols)[None, None, None, :]
    tau = rng.exponential(ds_ns * 1e-9, size=(batch, taps, 1, 1))
    pw = np.exp(-tau / (ds_ns * 1e-9))
    pw /= pw.sum(axis=1, keepdims=True)
    ph = rng.uniform(0, 2 * np.pi, size=(batch, taps, 1, 1))
    amp = np.sqrt(pw) * np.exp(1j * ph)
    nu = dop_hz * np.cos(rng.uniform(0, 2 * np.pi, size=(batch, taps, 1, 1)))
    df, ts = 15e3, 1e-3 / 14
    H = (amp * np.exp(-2j * np.pi * k * df * tau) * np.exp(2j * np.pi * nu * l * ts)).sum(axis=1)
    sc_idx = np.arange(cfg.pilot_scs) * (cfg.num_scs // cfg.pilot_scs)
    sym_idx = np.array([2, 11])[: cfg.pilot_symbols]
    sigma = 10 ** (-snr_db / 20)
    noise = (rng.standard_normal((batch, cfg.pilot_scs, cfg.pilot_symbols))
             + 1j * rng.standard_normal((batch, cfg.pilot_scs, cfg.pilot_symbols))) * sigma / math.sqrt(2)
    pilots = H[:, sc_idx][:, :, sym_idx] + noise
    return pilots.astype(np.complex64), H.astype(np.complex64)


# --------------------------------------------------------------------------------------
# "next" rows N2 / N3
# --------------------------------------------------------------------------------------
def extract_pilots(ls_grid: np.ndarray, pilot_size) -> np.ndarray:
    """Reference ``MatDataset._process_channel_data`` (src/data/dataset.py:118-139), batched: the non-zero entries of
    each sparse LS grid [B, scs, symbols] in row-major order, reshaped to [B, pilot_scs, pilot_symbols].
    Raises ValueError when a sample does not hold exactly pilot_scs * pilot_symbols non-zero entries."""
    ls_grid = np.asarray(ls_grid)
    expected = int(pilot_size[0]) * int(pilot_size[1])
    out = np.empty((ls_grid.shape[0], int(pilot_size[0]), int(pilot_size[1])), dtype=np.complex64)
    for b in range(ls_grid.shape[0]):
        flat = ls_grid[b].reshape(-1)
        nz = flat[flat != 0]
        if nz.size != expected:
            raise ValueError(f"Expected {expected} pilot values, got {nz.size}")
        out[b] = nz.reshape(out.shape[1:])
    return out


def linear_estimator(weight: np.ndarray, bias: np.ndarray, x: np.ndarray, ofdm_size) -> np.ndarray:
    """Reference ``LinearEstimator.forward`` (src/models/linear.py:82-95): flatten, W x + b, reshape."""
    y = x.reshape(x.shape[0], -1).astype(np.float64) @ weight.astype(np.float64).T + bias.astype(np.float64)
    return y.reshape(-1, int(ofdm_size[0]), int(ofdm_size[1]))
