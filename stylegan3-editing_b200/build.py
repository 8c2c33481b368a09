"""Build libsg3_b200.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

    python stylegan3-editing_b200/build.py [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  Objects go to csrc/build/ (git-ignored), the
shared library next to this file so it travels to the GPU box with the repository snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'libsg3_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC,-fvisibility=hidden', '-Xptxas', '-v',
          '--expt-relaxed-constexpr', '-I', os.path.join(os.path.dirname(HERE), 'include')]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(os.path.dirname(HERE), 'include', 'sg3_b200.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    path = os.path.join(CSRC, src)
    newest = max(os.path.getmtime(path), _deps_mtime())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, False
    cmd = [NVCC] + ARCH + CFLAGS + ['-c', path, '-o', obj]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(obj + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{res.stdout}')
    if verbose:
        print(res.stdout)
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in results]
    if force or any(changed for _, changed in results) or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-cudart', 'static']
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
