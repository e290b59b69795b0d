"""Builds libmagphase_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libmagphase_b200.so')
OBJ_DIR = os.path.join(HERE, 'csrc', 'build')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


# The kernel files are compiled with --use_fast_math: flush-to-zero, approximate float32 division / square root and fused
# multiply-add for FLOAT32 arithmetic only (3-6 % on the float32 kernels; float64 code -- analysis butterflies,
# Griffin-Lim, K-slice sums, cosine matrices -- is not affected).  The parity tests (1e-5 RMS) are the guard.
EXTRA_FLAGS = {f: ['--use_fast_math'] for f in ('mpb_synth_comp.cu', 'mpb_synthesis.cu', 'mpb_unwarp.cu', 'mpb_mel.cu', 'mpb_analysis.cu',
                                                          'mpb_mel_unwarp_tc.cu')}


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdr.append(os.path.join(HERE, '..', 'include', 'magphase_b200.h'))
    return hdr


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    hdrs = _deps()
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + '.o')
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [_nvcc()] + NVCC_FLAGS + EXTRA_FLAGS.get(os.path.basename(s), []) + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(8, len(jobs)))) as ex:
        failed = []
        for s, r in ex.map(compile_one, jobs):
            with open(os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + '.ptxas.log'), 'w') as f:
                f.write(r.stderr)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                failed.append(s)
        if failed:
            raise RuntimeError('nvcc failed on %s' % ', '.join(failed))
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + '.o') for s in srcs]
    if force or jobs or _stale(OUT, objs):
        cmd = [_nvcc(), '-shared', '-o', OUT] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
