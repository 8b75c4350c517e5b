"""Build the in-tree CUDA extension ``libupcp_cuda.so`` for sm_100a with nvcc.

Used by ``__graft_entry__.build()``.  The shared library is written next to this
file so that it travels with the repo snapshot to the GPU box (it is git-ignored,
not gpurun-ignored).  There is no JIT and no CPU fallback: if the library is
missing the package refuses to run (see ``_lib.py``).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libupcp_cuda.so')
OBJ_DIR = os.path.join(HERE, 'csrc', '_obj')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    # bit-exact fp64/fp32 arithmetic: no FMA contraction, IEEE div/sqrt, no FTZ
    '-fmad=false', '--prec-div=true', '--prec-sqrt=true', '--ftz=false',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-ffp-contract=off',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(path, deps):
    h = hashlib.sha256()
    h.update(' '.join(NVCC_FLAGS).encode())
    for p in [path] + deps:
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def build(verbose=False, force=False):
    """Compile every ``csrc/*.cu`` and link ``libupcp_cuda.so``.  Incremental."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                     if f.endswith(('.cuh', '.h')))
    headers.append(os.path.join(HERE, '..', 'include', 'upcp_cuda.h'))
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        stamp = obj + '.sha'
        dig = _digest(src, headers)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp)
                and open(stamp).read() == dig):
            continue
        jobs.append((src, obj, stamp, dig))

    def compile_one(job):
        src, obj, stamp, dig = job
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{res.stdout}\n{res.stderr}')
        if verbose:
            sys.stderr.write(res.stderr)
        with open(stamp, 'w') as f:
            f.write(dig)
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-lcudart']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    return LIB_PATH


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
