"""Build libqgb200.so in-tree with nvcc for sm_100a:  python -m pyqg_generative_b200.build"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'api.cu')
OUT = os.path.join(HERE, 'libqgb200.so')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-shared']


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'qgb200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    cmd = [nvcc] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUT, SRC]
    print(' '.join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='-v' in sys.argv)
