"""
CPU ORACLE — TEST INFRASTRUCTURE ONLY (see frei_oracle.py).

numpy restatement of the reference's wavelength binning: ``groupby_bins_agg`` with
``func=np.trapz`` (frei/interp.py:270-307) -> ``_binned_agg`` (:246-267) -> ``Trapz`` via
``AggregateTrapz._loop`` (:174-194, x = None so dx = 1).  ``np.add.at`` accumulates in array
order, i.e. the same sequence of additions as the reference's numba loop.
"""
import numpy as np
import pandas as pd


def groupby_bins_agg(array, group, bins, fill_value=0):
    array = np.asarray(array, dtype=np.float64)
    binned = pd.cut(np.ravel(group), bins)                    # interp.py:284
    codes = np.asarray(binned.codes).astype(np.int64)
    n_bins = binned.categories.size
    same = codes[:-1] == codes[1:]                            # interp.py:180
    if np.any(same & (codes[:-1] < 0)):
        raise ValueError("negative indices not supported")    # interp.py:182-183
    lead = array.shape[:-1]
    flat = array.reshape(-1, array.shape[-1])
    out = np.full((flat.shape[0], n_bins), float(fill_value))
    idx = np.flatnonzero(same)
    vals = (flat[:, idx] + flat[:, idx + 1]) / 2              # interp.py:186, 191 (dx = 1)
    filled = np.zeros(n_bins, dtype=bool)
    filled[codes[idx]] = True
    out[:, filled] = 0.0
    for r in range(flat.shape[0]):
        np.add.at(out[r], codes[idx], vals[r])                # ret[ri] += val, interp.py:201-202
    centres = np.array([0.5 * (b.left + b.right) for b in binned.categories])   # interp.py:304-306
    return out.reshape(lead + (n_bins,)), centres


def binned_opacity_one(opacity, wavelength_um, src_T, src_P, temperatures, pressures_bar, wl_bins):
    """
    numpy/scipy restatement of one species of binned_opacity's groupies branch
    (frei/opacity.py:128-146): crop, groupby_bins_agg(trapz) * bin width * 1e-3, then
    ``.interp(method='nearest', fill_value='extrapolate')`` — scipy's interp1d per axis.
    Returns [temperature, pressure, wavelength].
    """
    from scipy.interpolate import interp1d
    wl = np.asarray(wavelength_um)
    keep = (wl > wl_bins.min()) & (wl < wl_bins.max())
    binned, centres = groupby_bins_agg(np.asarray(opacity)[..., keep], wl[keep], wl_bins)
    binned = binned * (wl_bins[1:] - wl_bins[:-1]) * 1e-3
    f = interp1d(src_T, binned, kind='nearest', axis=0, fill_value='extrapolate', assume_sorted=False)
    out = f(temperatures)
    f = interp1d(src_P, out, kind='nearest', axis=1, fill_value='extrapolate', assume_sorted=False)
    return f(pressures_bar), centres
