"""Stub of matplotlib (only `use` is touched by the reference test conftest)."""


def use(*a, **k):
    pass
