"""Minimal stand-in for PyWavelets -- TEST INFRASTRUCTURE ONLY.

PyWavelets is an un-vendored, un-pinned dependency of the reference
(/root/reference/pyproject.toml:33) that is absent from this image.  The reference needs it for
(1) filter taps ``pywt.Wavelet(name)`` (src/ptwt/_util.py:82) and (2) default levels
``pywt.dwt_max_level`` / ``pywt.dwtn_max_level`` (src/ptwt/conv_transform.py:131,
conv_transform_2.py:138, conv_transform_3.py:117-119); neither is on the per-sample path.
With this shim (and the ``more_itertools`` one next to it) on ``sys.path`` the UNMODIFIED reference
imports from /root/reference/src and its four hot-path families run on CPU, which is how
``oracle/make_golden.py`` produces the committed fixtures.  Taps come from the same generated
table the product uses, so tap provenance can never cause a parity difference.
"""
from __future__ import annotations

from typing import Sequence

from pytorch_wavelet_toolbox_b200._wavelets import BuiltinWavelet, builtin_wavelet
from pytorch_wavelet_toolbox_b200._wavelets import dwt_max_level as _dml

from . import _functions  # noqa: F401

__version__ = "0.0-shim"


class Wavelet:
    def __init__(self, name: str = "custom", filter_bank=None):
        self.name = name
        if filter_bank is None:
            w = builtin_wavelet(name)
            bank = w.filter_bank
        else:
            bank = filter_bank.filter_bank if hasattr(filter_bank, "filter_bank") else filter_bank
        self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi = (list(map(float, f)) for f in bank)
        self.dec_len = len(self.dec_lo)
        self.rec_len = len(self.rec_lo)

    @property
    def filter_bank(self):
        return (self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi)

    def __len__(self):
        return self.dec_len


class ContinuousWavelet:  # imported by the reference's cwt module only
    def __init__(self, *a, **k):
        raise NotImplementedError("continuous wavelets are outside the hot path")


def DiscreteContinuousWavelet(name, filter_bank=None):
    return Wavelet(name, filter_bank)


def dwt_max_level(data_len: int, filter_len) -> int:
    if not isinstance(filter_len, int):
        filter_len = filter_len.dec_len
    return _dml(int(data_len), int(filter_len))


def dwtn_max_level(shape: Sequence[int], wavelet, axes=None) -> int:
    if isinstance(wavelet, str):
        wavelet = Wavelet(wavelet)
    flen = wavelet.dec_len if hasattr(wavelet, "dec_len") else len(wavelet)
    return min(dwt_max_level(int(n), flen) for n in shape)


def swt_max_level(input_len: int) -> int:
    lvl = 0
    while input_len > 0 and input_len % 2 == 0:
        input_len //= 2
        lvl += 1
    return lvl
