"""Configuration schemas accepted by the estimators: field-for-field compatible with the reference's
``src/config/schemas.py`` (:20-45 ``SystemConfig``, :48-110 device validation, :113-175 ``ModelConfig``)
and ``src/config/config_loader.py`` (:16-78), so the reference ``config/*.yaml`` files load unchanged.

Only the *contract* is restated (same field names, defaults, ``extra='forbid'``, same error types); there
is nothing to accelerate here.
"""
from __future__ import annotations

import logging
from pathlib import Path
from typing import List, Literal, Optional, Tuple, Union

import torch
import yaml
from pydantic import BaseModel, ConfigDict, Field, ValidationError, model_validator

_log = logging.getLogger(__name__)


class OFDMParams(BaseModel):
    num_scs: int = Field(..., gt=0, description="Number of OFDM subcarriers")
    num_symbols: int = Field(..., gt=0, description="Number of OFDM symbols")


class PilotParams(BaseModel):
    num_scs: int = Field(..., gt=0, description="Number of pilots across sub-carriers")
    num_symbols: int = Field(..., gt=0, description="Number of pilots across OFDM symbols")


class SystemConfig(BaseModel):
    """OFDM grid + pilot grid; the pilot grid may not exceed the OFDM grid (schemas.py:29-43)."""

    model_config = ConfigDict(extra="forbid")
    ofdm: OFDMParams
    pilot: PilotParams

    @model_validator(mode="after")
    def _pilots_fit(self):
        for axis, what in (("num_scs", "sub-carriers"), ("num_symbols", "symbols")):
            p, o = getattr(self.pilot, axis), getattr(self.ofdm, axis)
            if p > o:
                raise ValueError(f"Pilot {what} ({p}) cannot exceed OFDM {what} ({o})")
        return self


def _resolve_device(name: str) -> str:
    """Device-string validation with the reference's accept/reject behaviour (schemas.py:53-110)."""
    low = name.lower()
    has_mps = hasattr(torch.backends, "mps") and torch.backends.mps.is_available()
    if low == "auto":
        return "cuda" if torch.cuda.is_available() else ("mps" if has_mps else "cpu")
    if low == "cpu":
        return name
    if low.startswith("cuda"):
        if not torch.cuda.is_available():
            raise ValueError("CUDA is not available on this system")
        if ":" in low:
            try:
                index = int(low.split(":")[1])
            except ValueError:
                raise ValueError(f"Invalid CUDA device format: {low}") from None
            if index >= torch.cuda.device_count():
                raise ValueError(f"CUDA device {index} not available. "
                                 f"Available CUDA devices: {list(range(torch.cuda.device_count()))}")
        return name
    if low == "mps":
        if not has_mps:
            raise ValueError("MPS is not available/detected on this system")
        return name
    known = ["cpu"]
    if torch.cuda.is_available():
        known += ["cuda"] + [f"cuda:{i}" for i in range(torch.cuda.device_count())]
    if has_mps:
        known.append("mps")
    raise ValueError(f"Unsupported device: '{name}'. Available devices: {known}")


class BaseConfig(BaseModel):
    device: str = Field(default="cpu", description="Computing device to use")

    @model_validator(mode="after")
    def _check_device(self):
        self.device = _resolve_device(self.device)
        return self


class ModelConfig(BaseConfig):
    """Architecture parameters (schemas.py:113-175).  ``adafortitran`` requires the two adapter fields,
    ``fortitran`` forbids them, ``linear`` ignores them."""

    model_config = ConfigDict(extra="forbid")
    model_type: Literal["linear", "fortitran", "adafortitran"] = "fortitran"
    patch_size: Tuple[int, int] = Field(..., description="(subcarriers_per_patch, symbols_per_patch)")
    num_layers: int = Field(..., gt=0)
    model_dim: int = Field(..., gt=0)
    num_head: int = Field(..., gt=0)
    activation: Literal["relu", "gelu"] = "gelu"
    dropout: float = Field(default=0.1, ge=0.0, le=1.0)
    max_seq_len: int = Field(default=512, gt=0)
    pos_encoding_type: Literal["learnable", "sinusoidal"] = "learnable"
    adaptive_token_length: Optional[int] = Field(default=None, gt=0)
    channel_adaptivity_hidden_sizes: Optional[List[int]] = None

    @model_validator(mode="after")
    def _adapter_fields(self):
        adapter = {"channel_adaptivity_hidden_sizes": self.channel_adaptivity_hidden_sizes,
                   "adaptive_token_length": self.adaptive_token_length}
        if self.model_type == "adafortitran":
            for key, val in adapter.items():
                if val is None:
                    raise ValueError(f"{key} is required for AdaFortiTran model")
        elif self.model_type == "fortitran":
            for key, val in adapter.items():
                if val is not None:
                    raise ValueError(f"{key} should not be provided for FortiTran model")
        return self


def _read_yaml(path: Path, what: str) -> dict:
    if not path.exists():
        raise FileNotFoundError(f"{what} configuration file not found: {path}")
    if path.suffix != ".yaml":
        raise ValueError(f"{what} configuration file must be a .yaml file: {path}")
    try:
        with open(path, "r") as fh:
            return yaml.safe_load(fh)
    except yaml.YAMLError as exc:
        raise ValueError(f"Failed to parse YAML file {path}: {exc}")


class ConfigLoader:
    """YAML -> validated ``(SystemConfig, ModelConfig)``; pydantic errors surface as ``ValueError``
    (config_loader.py:57-58,69-70)."""

    def load_and_validate(self, system_config_path: Union[str, Path],
                          model_config_path: Union[str, Path]) -> Tuple[SystemConfig, ModelConfig]:
        sys_path, model_path = Path(system_config_path), Path(model_config_path)
        raw = []
        for path, what in ((sys_path, "System"), (model_path, "Model")):
            raw.append(_read_yaml(path, what))
        try:
            system = SystemConfig(**raw[0])
        except ValidationError as exc:
            raise ValueError(f"System configuration validation for {sys_path} failed:\n{exc}")
        try:
            model = ModelConfig(**raw[1])
        except ValidationError as exc:
            raise ValueError(f"Model configuration validation for {model_path} failed:\n{exc}")
        _log.info("loaded %s and %s", sys_path, model_path)
        return system, model


def load_config(system_config_path: Union[str, Path], model_config_path: Union[str, Path]):
    return ConfigLoader().load_and_validate(system_config_path, model_config_path)
