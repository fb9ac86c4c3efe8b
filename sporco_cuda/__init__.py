"""``sporco_cuda``-compatible package backed by libspcsc (sporco_b200).

The reference's ``sporco.cuda`` is an import shim: ``from sporco_cuda.util import *`` and
``from sporco_cuda.cbpdn import *`` (sporco/cuda/__init__.py:6-18).  With this directory on
``PYTHONPATH`` that import succeeds, ``sporco.cuda.have_cuda`` becomes True and the call sites
in the reference (``sporco/dictlrn/onlinecdl.py:161-165, 251-258, 270-275`` and the
``*_cuda.py`` example scripts) run on the B200 kernels without any edit of reference code.
Function signatures follow docs/source/modules/sporco.cuda.rst:60-251.
"""

__version__ = '0.0.10.b200'
