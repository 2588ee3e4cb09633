"""Placeholder for ``pywt._functions`` (imported by the reference's cwt module, off the hot path)."""


def scale2frequency(*args, **kwargs):
    raise NotImplementedError("continuous wavelets are outside the hot path")
