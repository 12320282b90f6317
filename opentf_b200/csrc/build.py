"""Build libntf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m opentf_b200.csrc.build [--force] [--verbose]

One object per .cu (compiled in parallel, skipped when up to date), then one shared library with a static
CUDA runtime so that it dlopen()s on a box without libcuda (symbol checks on the CPU-only CI box).
"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, 'libntf_b200.so')
OBJ = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '-Xcompiler', '-fvisibility=default', '--expt-relaxed-constexpr', '-I', os.path.join(ROOT, 'include')]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith('.cu'))


def _deps_mtime():
    hs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cuh', '.h'))] + [os.path.join(ROOT, 'include', 'ntf_b200.h')]
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    spath = os.path.join(HERE, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(spath), _deps_mtime()):
        return obj, ''
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', spath, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in res]
    if verbose:
        for (_, log), s in zip(res, srcs):
            if log: print(f'--- {s}\n{log}')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static', '-o', LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
