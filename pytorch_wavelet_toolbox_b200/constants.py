"""Public types of the hot path, mirroring ``ptwt.constants``.

The container types are re-stated (not imported) so the package works without ptwt
installed; they are structurally identical to the reference's
(``/root/reference/src/ptwt/constants.py:27-253``): same field names, same ordering,
plain ``tuple`` / ``dict`` / ``list`` subclasses, so coefficients produced here can be
fed to ptwt's own ``waverec*`` and vice versa.
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import Literal, NamedTuple, Protocol, Union

import torch

__all__ = [
    "SUPPORTED_DTYPES",
    "BoundaryMode",
    "OrthogonalizeMethod",
    "Wavelet",
    "WaveletTensorTuple",
    "WaveletDetailTuple2d",
    "WaveletDetailDict",
    "WaveletCoeff1d",
    "WaveletCoeff2d",
    "WaveletCoeffNd",
]

#: dtypes the transforms accept (reference constants.py:27); anything else -> ValueError.
SUPPORTED_DTYPES = {torch.float32, torch.float64}

BoundaryMode = Literal["constant", "zero", "reflect", "periodic", "symmetric"]
OrthogonalizeMethod = Literal["qr", "gramschmidt"]


class Wavelet(Protocol):
    """Duck type of a PyWavelets wavelet (reference constants.py:30-46)."""

    name: str
    dec_lo: Sequence[float]
    dec_hi: Sequence[float]
    rec_lo: Sequence[float]
    rec_hi: Sequence[float]
    dec_len: int
    rec_len: int
    filter_bank: tuple[Sequence[float], Sequence[float], Sequence[float], Sequence[float]]

    def __len__(self) -> int:  # pragma: no cover - protocol
        return len(self.dec_lo)


class WaveletTensorTuple(NamedTuple):
    """Filter bank given as four tensors (reference constants.py:49-82)."""

    dec_lo: torch.Tensor
    dec_hi: torch.Tensor
    rec_lo: torch.Tensor
    rec_hi: torch.Tensor

    @property
    def dec_len(self) -> int:
        return len(self.dec_lo)

    @property
    def rec_len(self) -> int:
        return len(self.rec_lo)

    @property
    def filter_bank(self):
        return self

    @classmethod
    def from_wavelet(cls, wavelet: "Wavelet", dtype: torch.dtype) -> "WaveletTensorTuple":
        return cls(*(torch.tensor(list(f), dtype=dtype) for f in (
            wavelet.dec_lo, wavelet.dec_hi, wavelet.rec_lo, wavelet.rec_hi)))


class WaveletDetailTuple2d(NamedTuple):
    """(H, V, D) detail bands of one 2-D level (reference constants.py:166-181)."""

    horizontal: torch.Tensor
    vertical: torch.Tensor
    diagonal: torch.Tensor


WaveletDetailDict = dict  # dict[str, torch.Tensor], keys like "aad" (reference constants.py:184)
WaveletCoeff1d = Sequence  # [cA_n, cD_n, ..., cD_1]
WaveletCoeff2d = tuple  # (cA_n, T_n, ..., T_1)
WaveletCoeffNd = tuple  # (cA_n, D_n, ..., D_1)

#: detail keys of one 3-D level in sub-band order k = 1..7 (reference conv_transform_3.py:131-141)
DETAIL_KEYS_3D = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")
