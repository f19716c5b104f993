"""CPU port of the reference forward over the reference's own third-party dependency (PyTorch).

TEST / BASELINE INFRASTRUCTURE ONLY (same rules as ``aft_oracle.py``): imported by ``tests/`` and by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, never by the product package.

Why it exists: the reference's arithmetic lives entirely inside an un-vendored, un-pinned dependency
(``requirements.txt:1``: ``torch``; installed here: torch 2.11.0+cu128), and the reference tree itself cannot
travel to the GPU box.  To time "the reference's CPU path" there, this module restates the composition of
``src/models/fortitran.py:184-233`` over the same torch operators the reference calls -- ``nn.Linear``
(fortitran.py:86), ``nn.Conv2d`` (enhancers.py:13-19), ``nn.Unfold``/``nn.Fold`` (patch_processors.py:22,53),
``nn.TransformerEncoder(nn.TransformerEncoderLayer(d, nhead, 2d, gelu, batch_first=True))`` (encoders.py:44-55)
and ``torch.complex`` (fortitran.py:180) -- so the CPU executes the same ATen kernels (including the
encoder fast path) as the reference would.  It is pinned against the golden vectors in
``tests/test_oracle_golden.py`` (bit-level agreement is expected; the gate is 1e-6).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
from torch import nn


class TorchPort(nn.Module):
    def __init__(self, sd: Dict[str, np.ndarray], adaptive: bool = True, num_layers: int = 6, d: int = 128, nhead: int = 4,
                 activation: str = "gelu", grid=(120, 14), patch=(3, 2)) -> None:
        super().__init__()
        t = {k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}
        self.adaptive, self.grid, self.patch = adaptive, grid, patch

        def lin(prefix):
            w, b = t[prefix + ".weight"], t[prefix + ".bias"]
            m = nn.Linear(w.shape[1], w.shape[0])
            m.weight.data.copy_(w), m.bias.data.copy_(b)
            return m

        def convs(prefix):
            mods = []
            for i, idx in enumerate((0, 2, 4, 6)):
                w = t[f"{prefix}.conv_block.{idx}.weight"]
                c = nn.Conv2d(w.shape[1], w.shape[0], 3, padding=1)
                c.weight.data.copy_(w), c.bias.data.copy_(t[f"{prefix}.conv_block.{idx}.bias"])
                mods.append(c)
                if i < 3:
                    mods.append(nn.ReLU())
            return nn.Sequential(*mods)

        self.up = lin("pilot_upsampler")
        self.enh = convs("initial_enhancer")
        self.ref = convs("final_refiner")
        if adaptive:
            self.mlps = nn.ModuleList(
                nn.Sequential(lin(f"channel_adapter.{n}_encoder.0"), nn.ReLU(), lin(f"channel_adapter.{n}_encoder.2"),
                              nn.ReLU(), lin(f"channel_adapter.{n}_encoder.4")) for n in ("snr", "ds", "dop"))
        self.l1 = lin("transformer_encoder.linear_1")
        self.l2 = lin("transformer_encoder.linear_2")
        pos_key = [k for k in t if k.endswith("position_embeddings") or k.endswith("positional_encoding.pe")][0]
        self.register_buffer("pos", t[pos_key].clone())
        layer = nn.TransformerEncoderLayer(d_model=d, nhead=nhead, dim_feedforward=2 * d, activation=activation,
                                           dropout=0.1, batch_first=True)
        self.enc = nn.TransformerEncoder(layer, num_layers=num_layers)
        own = self.enc.state_dict()
        for k in own:
            own[k] = t["transformer_encoder.transformer." + k]
        self.enc.load_state_dict(own)
        self.unfold = nn.Unfold(kernel_size=patch, stride=patch)
        self.fold = nn.Fold(output_size=grid, kernel_size=patch, stride=patch)
        self.eval()

    def _real(self, x: torch.Tensor, cond) -> torch.Tensor:
        n = x.shape[0]
        img = self.up(x.reshape(n, -1)).view(n, 1, *self.grid)
        enh = self.enh(img).squeeze(1)
        tok = self.unfold(enh.unsqueeze(1)).permute(0, 2, 1)
        if self.adaptive:
            z = [m(c).reshape(n, -1, 2) for m, c in zip(self.mlps, cond)]
            tok = torch.cat([tok] + z, dim=2)
        h = self.l1(tok)
        h = h + self.pos[:, : h.shape[1], :]
        h = self.l2(self.enc(h))
        rec = self.fold(h.permute(0, 2, 1)).squeeze(1)
        return self.ref((enh + rec).unsqueeze(1)).squeeze(1)

    @torch.no_grad()
    def forward(self, pilots: torch.Tensor, snr: Optional[torch.Tensor] = None, ds: Optional[torch.Tensor] = None,
                dop: Optional[torch.Tensor] = None) -> torch.Tensor:
        cond = None
        if self.adaptive:
            if snr is None or ds is None or dop is None:
                raise ValueError("meta_data is required when channel adaptation is enabled")
            cond = [torch.as_tensor(v, dtype=torch.float32).reshape(-1, 1) for v in (snr, ds, dop)]
        pilots = torch.as_tensor(pilots)
        return torch.complex(self._real(pilots.real, cond), self._real(pilots.imag, cond))
