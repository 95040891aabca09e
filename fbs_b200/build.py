"""Build the sm_100a C-ABI shared library in-tree (fbs_b200/_lib/libfbs_b200.so).

nvcc cross-compiles without a GPU.  Each .cu is compiled to its own object (in parallel, rebuilt only when it or a
header changed) and the objects are linked into one .so.  The .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_DIR = os.path.join(_HERE, '_lib')
OBJ_DIR = os.path.join(LIB_DIR, 'obj')
LIB_PATH = os.path.join(LIB_DIR, 'libfbs_b200.so')
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return 'nvcc'


def _headers_mtime():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    deps.append(os.path.join(_HERE, '..', 'include', 'fbs_b200.h'))
    return max(os.path.getmtime(d) for d in deps)


def _stale(src, obj, hdr_t):
    return (not os.path.exists(obj)) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t)


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return _headers_mtime() > t or any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = _headers_mtime()
    jobs = []
    for s in sources():
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ_DIR, s[:-3] + '.o')
        if force or _stale(src, obj, hdr_t):
            cmd = [_nvcc(), *NVCC_FLAGS, '-c', src, '-o', obj]
            if verbose:
                cmd[1:1] = ['-Xptxas', '-v']
            jobs.append(cmd)

    def run(cmd):
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
        return res.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for log in ex.map(run, jobs):
            if verbose:
                print(log)
    objs = [os.path.join(OBJ_DIR, s[:-3] + '.o') for s in sources()]
    run([_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH, *objs])
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
