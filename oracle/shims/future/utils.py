"""Stub of future.utils (only the two names the reference imports)."""
PY2 = False


def with_metaclass(meta, *bases):
    class _Tmp(meta):
        __call__ = type.__call__
        __init__ = type.__init__

        def __new__(cls, name, this_bases, d):
            if this_bases is None:
                return type.__new__(cls, name, (), d)
            return meta(name, bases, d)
    return _Tmp('temporary_class', None, {})
