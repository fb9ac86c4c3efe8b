"""TEST INFRASTRUCTURE: build the kernel sources for the CPU emulation harness.

Compiles sporco_b200/csrc/*.cu with g++ -DSPCSC_EMU against tests/emu/cuda_emu.{h,cpp}
into tests/emu/_build/libspcsc_emu.so.  The result executes the *same kernel source* on
the CPU (fibres as CUDA threads) so tests can check kernel logic without a GPU.  It is
never loaded by the sporco_b200 package itself -- only by tests that pass its path
explicitly -- and nothing is ever timed on it.
"""

import concurrent.futures
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'sporco_b200', 'csrc')
BUILD = os.path.join(HERE, '_build')
LIB = os.path.join(BUILD, 'libspcsc_emu.so')
SIZES = (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)
FLAGS = ['-O1', '-g', '-std=c++17', '-fPIC', '-DSPCSC_EMU', '-x', 'c++', '-I', HERE,
         '-I', CSRC, '-I', os.path.join(ROOT, 'include'), '-Wno-unused-result']


def _hash():
    h = hashlib.sha256()
    for d in (CSRC, HERE):
        for n in sorted(os.listdir(d)):
            p = os.path.join(d, n)
            if os.path.isfile(p) and not n.endswith('.py'):
                h.update(n.encode())
                h.update(open(p, 'rb').read())
    h.update(open(os.path.join(ROOT, 'include', 'spcsc.h'), 'rb').read())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('command failed: %s\n%s' % (' '.join(cmd), r.stdout))


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, 'stamp')
    digest = _hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    tasks = []
    for n in SIZES:
        obj = os.path.join(BUILD, 'size_%d.o' % n)
        tasks.append((obj, ['g++'] + FLAGS + ['-DSPCSC_SIZE=%d' % n, '-c',
                                             os.path.join(CSRC, 'size_inst.cu'), '-o', obj]))
    obj = os.path.join(BUILD, 'gen.o')
    tasks.append((obj, ['g++'] + FLAGS + ['-c', os.path.join(CSRC, 'gen_inst.cu'), '-o', obj]))
    obj = os.path.join(BUILD, 'spcsc.o')
    tasks.append((obj, ['g++'] + FLAGS + ['-c', os.path.join(CSRC, 'spcsc.cu'), '-o', obj]))
    obj = os.path.join(BUILD, 'cuda_emu.o')
    tasks.append((obj, ['g++'] + FLAGS + ['-c', os.path.join(HERE, 'cuda_emu.cpp'), '-o', obj]))
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(lambda t: _run(t[1]), tasks))
    _run(['g++', '-shared', '-o', LIB] + [t[0] for t in tasks] + ['-lpthread'])
    open(stamp, 'w').write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
