"""Option dictionaries with a closed key set.

Host-side mirror of ``sporco.cdict.ConstrainedDict`` (sporco/cdict.py:55-253): a dict whose
allowed keys -- including those of nested dicts -- are fixed by the class attribute
``defaults``; entries can be addressed with tuple keys (``opt['AutoRho', 'Period']``);
unknown keys raise :class:`UnknownKeyError`, replacing a sub-dict by a non-dict raises
:class:`InvalidValueError`.
"""

import pprint


class UnknownKeyError(KeyError):
    """Key that does not appear in the ``defaults`` tree (sporco/cdict.py:19-33)."""

    def __str__(self):
        k = self.args[0]
        return 'Unknown dictionary key: ' + ('.'.join(map(str, k))
                                             if isinstance(k, (list, tuple)) else str(k))

    __repr__ = __str__


class InvalidValueError(ValueError):
    """Non-dict value given where the ``defaults`` tree has a dict (sporco/cdict.py:37-51)."""

    def __str__(self):
        k = self.args[0]
        return 'Invalid dictionary value for key: ' + ('.'.join(map(str, k))
                                                       if isinstance(k, (list, tuple)) else str(k))

    __repr__ = __str__


def _walk(tree, path):
    node = tree
    for key in path:
        if not isinstance(node, dict):
            raise InvalidValueError(node)
        if key not in node:
            raise UnknownKeyError(tuple(path))
        node = dict.__getitem__(node, key)
    return node


class ConstrainedDict(dict):
    """dict restricted to the keys of ``defaults`` and pre-filled with them."""

    defaults = {}

    def __init__(self, d=None, pth=(), dflt=None):
        dict.__init__(self)
        self.pth = tuple(pth)
        self.dflt = self.__class__.defaults if dflt is None else dflt
        self.update(_walk(self.dflt, self.pth))
        self.update({} if d is None else d)

    # -- tree helpers kept under the reference's names
    @staticmethod
    def getparent(d, pth):
        return _walk(d, tuple(pth)[:-1])

    @staticmethod
    def getnode(d, pth):
        return _walk(d, tuple(pth))

    def check(self, key, value):
        allowed = _walk(self.dflt, self.pth)
        if key not in allowed:
            raise UnknownKeyError(self.pth + (key,))
        if isinstance(allowed[key], dict) and not isinstance(value, dict):
            raise InvalidValueError(self.pth + (key,))

    def update(self, d):
        for key in list(d.keys()):
            self[key] = d[key]

    def _locate(self, key):
        if isinstance(key, tuple):
            return _walk(self, key[:-1]), key[-1]
        return self, key

    def __setitem__(self, key, value):
        node, last = self._locate(key)
        plain = isinstance(value, dict) and not isinstance(value, ConstrainedDict)
        if plain and last in node:
            dict.__getitem__(node, last).update(value)
            return
        if plain:
            value = ConstrainedDict(value, node.pth + (last,), self.dflt)
        node.check(last, value)
        dict.__setitem__(node, last, value)

    def __getitem__(self, key):
        node, last = self._locate(key)
        if last not in node:
            raise UnknownKeyError(key)
        return dict.__getitem__(node, last)

    def __str__(self):
        return pprint.pformat(self)

    def __reduce__(self):
        return (_rebuild, (self.__class__, _plain(self), self.pth,
                           None if self.pth == () else self.dflt))


def _plain(d):
    return {k: (_plain(v) if isinstance(v, dict) else v) for k, v in dict.items(d)}


def _rebuild(cls, content, pth, dflt):
    return cls(content) if pth == () else ConstrainedDict(content, pth, dflt)


def keycmp(a, b, pth=()):
    """Raise if dict tree `b` has keys absent from `a` (sporco/cdict.py:310-347)."""
    for key in b:
        if key not in a:
            raise UnknownKeyError(pth + (key,))
        if isinstance(a[key], dict):
            if not isinstance(b[key], dict):
                raise InvalidValueError(pth + (key,))
            keycmp(a[key], b[key], pth + (key,))
