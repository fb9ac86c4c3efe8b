"""Build libspcsc.so (the CUDA extension) in-tree with nvcc for sm_100a.

    python -m sporco_b200.build [--force] [--jobs N]

Every transform length is its own translation unit (csrc/size_inst.cu with
-DSPCSC_SIZE=n) so the library compiles in parallel.  Objects go to sporco_b200/_build/,
the shared library to sporco_b200/libspcsc.so (git-ignored, shipped to the GPU box by
gpurun because it lives in the tree).
"""

import argparse
import concurrent.futures
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, '_build')
LIB = os.path.join(HERE, 'libspcsc.so')
SIZES = (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
              '-std=c++17', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
              '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def source_hash():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ['../../include/spcsc.h']
    for n in names:
        path = os.path.join(CSRC, n)
        if os.path.isfile(path):
            with open(path, 'rb') as f:
                h.update(n.encode())
                h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('command failed: %s\n%s' % (' '.join(cmd), r.stdout))
    return r.stdout


def build(force=False, jobs=None, verbose=False, extra_flags=()):
    """Compile and link libspcsc.so; returns its path.  No-op when sources are unchanged."""
    stamp = LIB + '.stamp'        # next to the library: objects need not travel with it
    digest = source_hash() + ' ' + ' '.join(extra_flags)
    if (not force and os.path.exists(LIB) and os.path.exists(stamp)
            and open(stamp).read() == digest):
        return LIB
    nvcc = find_nvcc()
    os.makedirs(BUILD, exist_ok=True)
    flags = NVCC_FLAGS + list(extra_flags)
    tasks = []
    for n in SIZES:
        obj = os.path.join(BUILD, 'size_%d.o' % n)
        tasks.append((obj, [nvcc] + flags + ['-DSPCSC_SIZE=%d' % n, '-c',
                                            os.path.join(CSRC, 'size_inst.cu'), '-o', obj]))
    obj = os.path.join(BUILD, 'gen.o')
    tasks.append((obj, [nvcc] + flags + ['-c', os.path.join(CSRC, 'gen_inst.cu'), '-o', obj]))
    obj = os.path.join(BUILD, 'spcsc.o')
    tasks.append((obj, [nvcc] + flags + ['-c', os.path.join(CSRC, 'spcsc.cu'), '-o', obj]))
    jobs = jobs or min(len(tasks), os.cpu_count() or 4)
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs) as ex:
        outs = list(ex.map(lambda t: _run(t[1]), tasks))
    if verbose:
        for o in outs:
            if o.strip():
                print(o)
    _run([nvcc, '-shared', '-o', LIB] + [t[0] for t in tasks] +
         ['-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC', '-ldl'])
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--jobs', type=int, default=None)
    ap.add_argument('--verbose', action='store_true')
    ap.add_argument('--ptxas-info', action='store_true', help='pass -Xptxas -v')
    a = ap.parse_args()
    extra = ('-Xptxas', '-v') if a.ptxas_info else ()
    print(build(force=a.force, jobs=a.jobs, verbose=a.verbose or a.ptxas_info, extra_flags=extra))
