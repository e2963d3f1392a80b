"""
Multi-GPU parity (wavelength sharding) under pytest: spawns one process per GPU with
torch.distributed.run on min(2, device_count) GPUs of this node and checks the sharded solve
against the single-GPU solve of the same problem inside every rank (tests/dist_worker.py):
same iteration count, spectrum, T history and dtaus to rounding of the summation order, for the
fused peer-memory exchange (collective='p2p') and the NCCL all-reduce, with uneven shards
(odd wavelength count), gather='all' and gather='local', and the sharded diagnostics.
Skipped on a single-GPU box.
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run_worker(world, extra):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'dist_worker.py')] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    log = res.stdout + res.stderr
    logdir = os.environ.get('FREI_DIST_LOGDIR')
    if logdir:
        os.makedirs(logdir, exist_ok=True)
        with open(os.path.join(logdir, f'dist_parity_n{world}_{"_".join(extra).replace("--", "")}.log'), 'w') as fh:
            fh.write(' '.join(cmd) + '\n' + log)
    assert res.returncode == 0, log[-3000:]
    assert log.count('-> OK') == 2 * world and 'MISMATCH' not in log, log[-3000:]
    return log


@pytest.mark.parametrize('collective', ['p2p', 'nccl'])
def test_sharded_solve_matches_single_gpu(collective):
    n = _n_gpus()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    log = _run_worker(2, ['--collective', collective])
    assert ('p2p-fused' if collective == 'p2p' else '[nccl]') in log


def test_sharded_solve_all_gpus_large_shapes():
    """All GPUs of the box, C3-like shape (8 species, 100 layers), automatic collective."""
    n = _n_gpus()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    _run_worker(n, ['--layers', '100', '--species', '8', '--nlam', '20002'])
