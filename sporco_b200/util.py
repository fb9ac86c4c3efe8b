"""Small host utilities: the labelled stopwatch the solvers expose as ``.timer``
(mirror of ``sporco.util.Timer``, sporco/util.py:574-806)."""

from timeit import default_timer


class Timer(object):
    """A set of independent, labelled, accumulating stopwatches."""

    def __init__(self, labels=None, dfltlbl='main', alllbl='all'):
        self.t0 = {}
        self.td = {}
        self.dfltlbl = dfltlbl
        self.alllbl = alllbl
        for lbl in self._as_list(labels, none_ok=True):
            self.td[lbl] = 0.0
            self.t0[lbl] = None

    def _as_list(self, labels, none_ok=False):
        if labels is None:
            return [] if none_ok else [self.dfltlbl]
        if isinstance(labels, (list, tuple)):
            return list(labels)
        if labels == self.alllbl:
            return list(self.t0.keys())
        return [labels]

    def start(self, labels=None):
        now = default_timer()
        for lbl in self._as_list(labels):
            if lbl not in self.td:
                self.td[lbl] = 0.0
                self.t0[lbl] = None
            if self.t0[lbl] is None:
                self.t0[lbl] = now

    def stop(self, labels=None):
        now = default_timer()
        for lbl in self._as_list(labels):
            if lbl not in self.t0:
                raise KeyError('Unrecognized timer key %s' % lbl)
            if self.t0[lbl] is not None:
                self.td[lbl] += now - self.t0[lbl]
                self.t0[lbl] = None

    def reset(self, labels=None):
        now = default_timer()
        for lbl in self._as_list(labels):
            if lbl not in self.t0:
                raise KeyError('Unrecognized timer key %s' % lbl)
            if self.t0[lbl] is not None:
                self.t0[lbl] = now
            self.td[lbl] = 0.0

    def elapsed(self, label=None, total=True):
        now = default_timer()
        lbl = self.dfltlbl if label is None else label
        if lbl not in self.t0:
            return 0.0
        run = 0.0 if self.t0[lbl] is None else now - self.t0[lbl]
        return self.td[lbl] + run if total else run

    def labels(self):
        return self.t0.keys()

    def __str__(self):
        return '\n'.join('%-16s %.2e s%s' % (lbl, self.elapsed(lbl),
                                            '' if self.t0[lbl] is None else ' (running)')
                         for lbl in sorted(self.t0))
