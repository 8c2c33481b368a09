"""Build libsg3_b200.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

    python stylegan3-editing_b200/build.py [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  Objects go to csrc/build/ (git-ignored), the
shared library next to this file so it travels to the GPU box with the repository snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# Tuning experiments: SG3_NVCC_EXTRA="-DSG3_FL_WARPS=1 ..." builds a variant library next to the default one
# (SG3_LIB_SUFFIX names it, e.g. "_w1" -> libsg3_b200_w1.so; load it with SG3_B200_LIB=<path>, see capi.py).
_EXTRA = os.environ.get('SG3_NVCC_EXTRA', '').split()
_SUFFIX = os.environ.get('SG3_LIB_SUFFIX', '')
OBJ = os.path.join(CSRC, 'build' + _SUFFIX)
LIB = os.path.join(HERE, f'libsg3_b200{_SUFFIX}.so')
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


def source_hash():
    """sha256 over the names and contents of every source the library is built from (csrc/*.{cu,cuh,h}, include/sg3_b200.h)."""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh', '.h'))]
    files.append(os.path.join(os.path.dirname(HERE), 'include', 'sg3_b200.h'))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    path = os.path.join(CSRC, src)
    newest = max(os.path.getmtime(path), _deps_mtime())
    extra = []
    if src == 'capi.cu':
        # the build-info string carries the source hash: this one file is rebuilt whenever any source changed
        digest = source_hash()
        extra = ['-DSG3_SOURCE_HASH=' + digest]
        stamp = obj + '.hash'
        if not (os.path.exists(stamp) and open(stamp).read() == digest):
            force = True
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, False
    cmd = [NVCC] + ARCH + CFLAGS + _EXTRA + extra + ['-c', path, '-o', obj]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(obj + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{res.stdout}')
    if verbose:
        print(res.stdout)
    if extra:
        with open(obj + '.hash', 'w') as f:
            f.write(extra[0].split('=', 1)[1])
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
