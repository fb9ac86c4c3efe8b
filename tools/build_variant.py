"""Build a variant of the library with extra nvcc flags (experiments / bisection only):
    python tools/build_variant.py NAME -DFLAG ...   ->  sporco_b200/libspcsc_NAME.so
Select it at run time with SPCSC_LIBRARY=<path> in the tools/ scripts."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sporco_b200 import build   # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
build.BUILD = os.path.join(build.HERE, '_build_' + name)
build.LIB = os.path.join(build.HERE, 'libspcsc_%s.so' % name)
print(build.build(force=True, extra_flags=tuple(flags)))
