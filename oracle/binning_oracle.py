"""
CPU ORACLE — TEST INFRASTRUCTURE ONLY (see frei_oracle.py).

numpy restatement of the reference's wavelength binning: ``groupby_bins_agg`` with
``func=np.trapz`` (frei/interp.py:270-307) -> ``_binned_agg`` (:246-267) -> ``Trapz`` via
``AggregateTrapz._loop`` (:174-194, x = None so dx = 1).  ``np.add.at`` accumulates in array
order, i.e. the same sequence of additions as the reference's numba loop.
"""
import numpy as np
import pandas as pd


def groupby_bins_agg(array, group, bins, fill_value=0, keep_dtype=False):
    """
    keep_dtype=True replays the reference's accumulator type: numpy_groupies' check_dtype keeps a
    floating input type, so float32 samples are summed into a float32 ``ret`` (every
    ``ret[ri] += val`` adds in float64 and rounds back to float32, frei/interp.py:57, 186-202).
    Default: float64 accumulation (what the GPU path does for either input type).
    """
    if keep_dtype and np.asarray(array).dtype == np.float32:
        return _groupby_bins_agg_f32(np.asarray(array), group, bins, fill_value)
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


def _groupby_bins_agg_f32(array, group, bins, fill_value):
    binned = pd.cut(np.ravel(group), bins)
    codes = np.asarray(binned.codes).astype(np.int64)
    n_bins = binned.categories.size
    lead = array.shape[:-1]
    flat = array.reshape(-1, array.shape[-1])
    out = np.full((flat.shape[0], n_bins), fill_value, dtype=np.float32)
    for r in range(flat.shape[0]):
        for i in range(len(codes) - 1):
            if codes[i] == codes[i + 1]:                      # interp.py:180
                if codes[i] < 0:
                    raise ValueError("negative indices not supported")
                avg_y = np.float64(flat[r, i] + flat[r, i + 1]) / 2          # float32 sum, then / 2
                out[r, codes[i]] = np.float32(np.float64(out[r, codes[i]]) + avg_y)
    centres = np.array([0.5 * (b.left + b.right) for b in binned.categories])
    return out.reshape(lead + (n_bins,)), centres


def binned_opacity_exact_one(opacity, wavelength_um, src_T, src_P, temperatures, pressures_bar, wl_bins, lam):
    """
    numpy/scipy restatement of one species of binned_opacity's groupies=False branch
    (frei/opacity.py:150-167 with mapfunc_exact, :29-40): per non-empty pandas.cut bin, nearest
    (T, P) lookup, np.trapz over the bin's samples divided by their wavelength span, at their
    mean wavelength; then linear interpolation with extrapolation onto ``lam``.
    Returns [wavelength, temperature, pressure] (the reference's dimension order).
    """
    from scipy.interpolate import interp1d
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    wl = np.asarray(wavelength_um, dtype=np.float64)
    codes = np.asarray(pd.cut(wl, wl_bins).codes)
    f = interp1d(src_T, np.asarray(opacity, dtype=np.float64), kind='nearest', axis=0,
                 fill_value='extrapolate', assume_sorted=False)
    op = f(temperatures)
    f = interp1d(src_P, op, kind='nearest', axis=1, fill_value='extrapolate', assume_sorted=False)
    op = f(pressures_bar)                                     # [T', P', n]
    parts, xs = [], []
    with np.errstate(invalid='ignore', divide='ignore'):
        for b in range(len(wl_bins) - 1):
            idx = np.flatnonzero(codes == b)
            if idx.size == 0:
                continue                                      # xarray's groupby skips empty bins
            w = wl[idx]
            parts.append(trapz(op[..., idx], w, axis=-1) / (w.max() - w.min()))   # :38-40
            xs.append(w.mean())
    binned = np.stack(parts, axis=0)
    f = interp1d(np.array(xs), binned, kind='linear', axis=0, bounds_error=False,
                 fill_value='extrapolate', assume_sorted=False)                  # :163-166
    return f(np.asarray(lam, dtype=np.float64))


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
