"""In-tree build of librgl_b200.so (hand-written sm_100a kernels + the C ABI of include/rgl_b200.h).

    python -m relationalgraphlearning_b200.build

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'librgl_b200.so')

OBJ = os.path.join(PKG, 'csrc', '_obj')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(CSRC, '*.h')) + \
        [os.path.join(os.path.dirname(PKG), 'include', 'rgl_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    # one nvcc per translation unit, in parallel (no device code crosses a TU), then one link
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get('NVCC', 'nvcc')
    os.makedirs(OBJ, exist_ok=True)
    hdr = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(CSRC, '*.h')) + \
        [os.path.join(os.path.dirname(PKG), 'include', 'rgl_b200.h')]
    hdr_t = max(os.path.getmtime(h) for h in hdr)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, 0, ''
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return obj, res.returncode, res.stdout

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    for obj, rc, out in results:
        if verbose or rc:
            sys.stderr.write(out)
    if any(rc for _, rc, _ in results):
        raise RuntimeError('nvcc failed building librgl_b200.so')
    res = subprocess.run([nvcc, '-shared', '-o', LIB] + [o for o, _, _ in results],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode:
        sys.stderr.write(res.stdout)
        raise RuntimeError('link failed building librgl_b200.so')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
