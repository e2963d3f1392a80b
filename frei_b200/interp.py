"""
Wavelength binning of line-by-line opacities — same entry point as ``frei/interp.py``
(``groupby_bins_agg``), run on the GPU (``csrc/binning.cu``).

The reference bins ~1e6-1e8 line-by-line samples per (temperature, pressure) into the
Grid's wavelength bins at load time (``frei/opacity.py:137-139``): samples are assigned
to right-closed bins with ``pandas.cut`` and, per bin, consecutive samples that both fall
in the bin contribute ``(a_i + a_{i+1}) / 2`` (a trapezoid sum with unit spacing,
``frei/interp.py:174-194``).  Host code here only derives the bin code of each sample
(the job ``pandas.cut`` does in the reference) and the runs of equal codes.
"""
import numpy as np

from . import _cabi
from ._cabi import FREI_F32, FREI_F64

__all__ = ['groupby_bins_agg']


class BinnedArray(np.ndarray):
    """ndarray with the bin-centre coordinate the reference attaches (frei/interp.py:304-306)."""
    wavelength = None


def cut_codes(values, bins):
    """Bin code of every sample: ``pandas.cut(values, bins).codes`` (right-closed, -1 outside)."""
    values = np.asarray(values, dtype=np.float64)
    bins = np.asarray(bins, dtype=np.float64)
    codes = np.searchsorted(bins, values, side='left') - 1
    codes[(values <= bins[0]) | (values > bins[-1]) | np.isnan(values)] = -1
    return codes.astype(np.int64)


def runs_by_bin(codes, n_bins):
    """Runs of equal consecutive codes as CSR over bins: (run_start, run_end, bin_first_run)."""
    codes = np.asarray(codes)
    n = codes.shape[0]
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(n_bins + 1, np.int32)
    change = np.flatnonzero(codes[1:] != codes[:-1]) + 1
    starts = np.concatenate([[0], change])
    ends = np.concatenate([change, [n]])
    rcode = codes[starts]
    # the reference raises as soon as two consecutive samples share a negative code
    # (frei/interp.py:182-183)
    if np.any((rcode < 0) & (ends - starts >= 2)):
        raise ValueError("negative indices not supported")
    if np.any(rcode >= n_bins):
        raise ValueError("one or more indices in group_idx are too large")
    keep = (rcode >= 0) & (ends - starts >= 2)          # single samples contribute nothing
    starts, ends, rcode = starts[keep], ends[keep], rcode[keep]
    order = np.argsort(rcode, kind='stable')             # runs of a bin stay in array order
    starts, ends, rcode = starts[order], ends[order], rcode[order]
    first = np.searchsorted(rcode, np.arange(n_bins + 1), side='left').astype(np.int32)
    return starts.astype(np.int64), ends.astype(np.int64), first


def _device_samples(array):
    """Samples as a contiguous CUDA tensor [..., n] (fp32 stays fp32; everything else fp64)."""
    import torch
    dev = torch.device('cuda', torch.cuda.current_device())
    if torch.is_tensor(array):
        a = array.to(dev)
    else:
        a_np = np.asarray(getattr(array, 'values', array))
        if a_np.dtype not in (np.float32, np.float64):
            a_np = a_np.astype(np.float64)
        a = torch.from_numpy(np.ascontiguousarray(a_np)).to(dev)
    if a.dtype not in (torch.float32, torch.float64):
        a = a.double()
    return a.contiguous()


def bin_trapz_device(a, group, bins, positions=False):
    """
    Trapezoid sum of the samples ``a[..., n]`` (CUDA tensor) per bin of ``group[n]``: unit spacing
    (``positions=False``: the reference's Trapz aggregation, frei/interp.py:174-194) or with the
    sample positions ``group`` themselves as x (xarray's ``integrate``, frei/opacity.py:39).
    Returns ``(out[..., n_bins] fp64 CUDA tensor, first)`` with ``first`` the CSR run offsets per
    bin (``first[b] == first[b + 1]``: no pair of consecutive samples fell into bin b).
    """
    import torch
    lib = _cabi.load()
    _cabi.require_cuda()
    dev = a.device
    g = np.asarray(group, dtype=np.float64).ravel()
    n_bins = bins.shape[0] - 1
    lead = tuple(a.shape[:-1])
    n = a.shape[-1]
    if n != g.shape[0]:
        raise ValueError('array and group differ in length')
    rows = int(np.prod(lead)) if lead else 1
    a2 = a.reshape(rows, n)
    starts, ends, first = runs_by_bin(cut_codes(g, bins), n_bins)
    d_s = torch.from_numpy(np.ascontiguousarray(starts) if starts.size else np.zeros(1, np.int64)).to(dev)
    d_e = torch.from_numpy(np.ascontiguousarray(ends) if ends.size else np.zeros(1, np.int64)).to(dev)
    d_f = torch.from_numpy(first).to(dev)
    d_x = torch.from_numpy(g).to(dev) if positions else None
    out = torch.empty((rows, n_bins), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    for r0 in range(0, rows, 65535):
        r1 = min(rows, r0 + 65535)
        _cabi.check(lib.frei_b200_bin_trapz(
            a2[r0:r1].data_ptr(), FREI_F32 if a.dtype == torch.float32 else FREI_F64, _cabi.ptr(d_x),
            r1 - r0, n, n, d_s.data_ptr(), d_e.data_ptr(), d_f.data_ptr(), n_bins,
            out[r0:r1].data_ptr(), st))
    return out.reshape(lead + (n_bins,)), first


def groupby_bins_agg(array, group, bins, func='trapz', fill_value=0, dtype=None, **cut_kwargs):
    """
    Aggregate ``array[..., n_samples]`` over the bins of ``group[n_samples]`` — the reference's
    ``groupby_bins_agg`` (frei/interp.py:270-307) for its only call site, ``func=np.trapz`` /
    ``'trapz'``.  ``bins`` are the n_bins + 1 edges (``Grid.wl_bins``).  Returns
    ``[..., n_bins]`` with the bin centres in ``.wavelength``.  Accepts numpy arrays, torch CUDA
    tensors (no copy) or xarray-like objects (``.values``).  Sums are accumulated in fp64 (the
    reference accumulates float32 samples in float32, frei/interp.py:57, 201-202).
    """
    import torch
    if cut_kwargs:
        raise NotImplementedError('pandas.cut options other than the defaults are not used by frei')
    if not (func == 'trapz' or func is getattr(np, 'trapz', None) or func is getattr(np, 'trapezoid', None)):
        raise NotImplementedError("only func='trapz' is on frei's path (frei/opacity.py:137-139)")
    _cabi.require_cuda()
    g = np.asarray(getattr(group, 'values', group), dtype=np.float64).ravel()
    bins = np.asarray(getattr(bins, 'value', bins), dtype=np.float64)
    a = _device_samples(array)
    res, first = bin_trapz_device(a, g, bins)
    if fill_value != 0:
        empty = torch.from_numpy(first[1:] == first[:-1]).to(res.device)
        res[..., empty] = fill_value
    if torch.is_tensor(array):
        return res
    host = res.cpu().numpy()
    if dtype is not None:
        host = host.astype(dtype)
    host = host.view(BinnedArray)
    host.wavelength = bin_centres(bins)
    return host


def bin_centres(bins):
    """
    Bin-centre coordinate as the reference computes it (frei/interp.py:304-306): from the
    interval labels of ``pandas.cut``, whose edges are rounded to 3 significant decimals
    (pandas' default ``precision=3``) — not from the exact edges.
    """
    bins = np.asarray(bins, dtype=np.float64)
    try:
        import pandas as pd
        cats = pd.cut(bins[1:2], bins).categories
        return np.array([0.5 * (b.left + b.right) for b in cats])
    except ImportError:                                     # pragma: no cover
        return 0.5 * (bins[:-1] + bins[1:])
