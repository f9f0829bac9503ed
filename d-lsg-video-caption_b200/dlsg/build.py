"""Build libdlsg.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), 'csrc')
LIB = os.path.join(HERE, 'libdlsg.so')
SOURCES = ['api.cu', 'gemm_tc.cu', 'gemm_simt.cu', 'rowops.cu', 'decode_ops.cu', 'graph_ops.cu', 'fused_step.cu', 'norm_bf16.cu', 'region_agg.cu', 'lstm_step.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, 'common.cuh'), os.path.join(os.path.dirname(os.path.dirname(HERE)), 'include', 'dlsg.h')]
    objs = []
    jobs = []
    for s in srcs:
        o = s[:-3] + '.o'
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + ['-c', s, '-o', o], capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed for %s' % s)
    if jobs or force or not os.path.exists(LIB):
        r = subprocess.run([NVCC, '-shared', '-o', LIB] + objs + ['-cudart', 'static'], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
