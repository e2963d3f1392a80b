"""
In-tree build of the C-ABI library ``frei_b200/_lib/libfrei_b200.so`` with nvcc
for sm_100a.  The ``.so`` is git-ignored but travels with the working tree to
the GPU box, so nothing is compiled there.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, '_lib')
LIBNAME = 'libfrei_b200.so'
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-shared', '-Xcompiler', '-fPIC',
    '-Xptxas', '-v',
]


def _extra_flags():
    """Experiment knobs, e.g. FREI_B200_NVCC_EXTRA='-DSWEEP_MINB=6 -DSWEEP_THREADS=64'."""
    return os.environ.get('FREI_B200_NVCC_EXTRA', '').split()


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    hsh = hashlib.sha256()
    files = _sources() + sorted(
        os.path.join(d, f) for d in (CSRC, INCLUDE) for f in os.listdir(d)
        if f.endswith(('.h', '.cuh')))
    for f in files:
        with open(f, 'rb') as fh:
            hsh.update(fh.read())
    hsh.update(' '.join(NVCC_FLAGS + _extra_flags()).encode())
    return hsh.hexdigest()


def find_nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    return nvcc if os.path.exists(nvcc) else None


def build(force=False, verbose=False):
    """Compile if the sources changed since the last build.  Returns the library path."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'build.sha256')
    digest = _digest()
    if (not force and os.path.exists(lib_path()) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest):
        return lib_path()
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError('nvcc not found: cannot build libfrei_b200.so')
    cmd = [nvcc] + NVCC_FLAGS + _extra_flags() + ['-I', INCLUDE, '-o', lib_path()] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(LIBDIR, 'build.log'), 'w') as fh:
        fh.write(' '.join(cmd) + '\n' + log)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + log[-4000:])
    if verbose:
        print(log)
    with open(stamp, 'w') as fh:
        fh.write(digest)
    return lib_path()


def build_variant(tag, flags):
    """Experiment builds (A/B of compile-time knobs): ``_lib/variants/libfrei_b200_<tag>.so``,
    selected at run time with ``FREI_B200_LIB=<path>``.  Not used by the product path."""
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError('nvcc not found')
    vdir = os.path.join(LIBDIR, 'variants')
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, f'libfrei_b200_{tag}.so')
    cmd = [nvcc] + NVCC_FLAGS + list(flags) + ['-I', INCLUDE, '-o', out] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(out + '.log', 'w') as fh:
        fh.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + (res.stdout + res.stderr)[-4000:])
    return out


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--variant':
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force='--force' in sys.argv, verbose=True))
