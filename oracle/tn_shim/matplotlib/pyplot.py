"""Stub (see matplotlib/__init__.py)."""


def __getattr__(name):
    raise RuntimeError('matplotlib is not installed; plotting is out of scope for the oracle')
