"""Stand-in for ``more_itertools.grouper`` (reference src/ptwt/_util.py:15,804) -- TEST INFRASTRUCTURE ONLY."""
from itertools import zip_longest


def grouper(iterable, n, incomplete="fill", fillvalue=None):
    args = [iter(iterable)] * n
    return zip_longest(*args, fillvalue=fillvalue)
