"""Compile the plain-C oracle natives into oracle/_build/libupcp_oracle.so (gcc)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'native', 'upcp_oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB_PATH = os.path.join(OUT_DIR, 'libupcp_oracle.so')


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC)):
        return LIB_PATH
    gcc = shutil.which('gcc') or 'gcc'
    cmd = [gcc, '-O2', '-fopenmp', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared',
           '-std=c11', '-o', LIB_PATH, SRC, '-lm']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'gcc failed:\n{res.stdout}\n{res.stderr}')
    return LIB_PATH


if __name__ == '__main__':
    print(build(force=True))
