"""In-tree build of libdiffspectra_b200.so with nvcc for sm_100a (no torch extension machinery needed:
the boundary is a plain C-ABI shared library loaded with ctypes)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libdiffspectra_b200.so')
SOURCES = ['api_core.cu', 'api_sampling.cu', 'gemm_tc.cu', 'coord_head_tc.cu', 'edge_ffn_tc.cu', 'gemm_simt.cu', 'dmt_kernels.cu', 'sampler_kernels.cu',
           'specformer_kernels.cu', 'weights.cu']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '--threads', '0']


def _stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'diffspectra_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        cmd = [nvcc] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    fail = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('== %s ==\n%s\n' % (src, out))
        fail |= p.returncode != 0
    if fail:
        raise RuntimeError('nvcc failed')
    subprocess.check_call([nvcc, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(OUT)
