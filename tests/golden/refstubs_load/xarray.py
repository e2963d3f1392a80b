"""
Stand-in for the xarray features the reference's LOAD-TIME path uses (frei/interp.py:270-307,
frei/opacity.py:29-40, 128-167, 467-483), numpy-backed.  Only used by
tests/golden/run_reference_load.py; the semantics restated here are xarray's documented ones:

* ``DataArray.interp`` with 1-D indexers on several dimensions interpolates orthogonally, one
  dimension after the other, with ``scipy.interpolate.interp1d(kind=method, bounds_error=False,
  fill_value=..., assume_sorted=False)``;
* ``Dataset.groupby_bins(name, bins)`` cuts with ``pandas.cut`` (right-closed), skips empty bins,
  and ``.map(func)`` concatenates the per-group results along the grouped dimension in bin order;
* ``DataArray.integrate(dim)`` is the trapezoid rule over the coordinate of ``dim``;
* ``apply_ufunc`` with core dims moves them to the end and calls the function on the raw arrays.
"""
import numpy as np
import pandas as pd
from scipy.interpolate import interp1d


def _raw(x):
    if isinstance(x, DataArray):
        return x.values
    return np.asarray(getattr(x, 'value', x))


class _Coords(dict):
    def __init__(self, owner, *a, **k):
        super().__init__(*a, **k)
        self._owner = owner

    def __setitem__(self, key, val):
        super().__setitem__(key, np.asarray(val))
        if key not in self._owner.dims:
            raise KeyError(key)


class DataArray:
    __array_ufunc__ = None
    __array_priority__ = 20000

    def __init__(self, data, dims=None, coords=None, name=None):
        self.values = np.asarray(data)
        self.dims = (dims,) if isinstance(dims, str) else tuple(dims or ())
        self.coords = _Coords(self, {k: np.asarray(v) for k, v in (coords or {}).items()})
        self.name = name

    # -- basics ----------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def sizes(self):
        return dict(zip(self.dims, self.values.shape))

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return len(self.values)

    def __getattr__(self, name):
        coords = self.__dict__.get('coords', {})
        if name in coords:                                   # a coordinate as a 1-D DataArray
            return DataArray(coords[name], (name,), {name: coords[name]}, name=name)
        raise AttributeError(name)

    def copy(self, data=None):
        return DataArray(self.values.copy() if data is None else np.asarray(data), self.dims,
                         dict(self.coords), self.name)

    def rename(self, mapping):
        dims = tuple(mapping.get(d, d) for d in self.dims)
        coords = {mapping.get(k, k): v for k, v in self.coords.items()}
        return DataArray(self.values, dims, coords, self.name)

    def _binary(self, other, op):
        return DataArray(op(self.values, _raw(other)), self.dims, dict(self.coords), self.name)

    def __gt__(self, o): return self._binary(o, np.greater)
    def __lt__(self, o): return self._binary(o, np.less)
    def __and__(self, o): return self._binary(o, np.logical_and)
    def __mul__(self, o): return self._binary(o, np.multiply)
    __rmul__ = __mul__
    def __truediv__(self, o): return self._binary(o, np.true_divide)
    def __sub__(self, o): return self._binary(o, np.subtract)

    def max(self): return DataArray(self.values.max())
    def min(self): return DataArray(self.values.min())
    def mean(self): return DataArray(self.values.mean())

    # -- selection -------------------------------------------------------------------------------
    def where(self, cond, drop=False):
        assert drop and len(cond.dims) == 1
        dim = cond.dims[0]
        ax = self.dims.index(dim)
        keep = np.flatnonzero(cond.values)
        coords = dict(self.coords)
        coords[dim] = coords[dim][keep]
        return DataArray(np.take(self.values, keep, axis=ax), self.dims, coords, self.name)

    def isel_dim(self, dim, idx):
        ax = self.dims.index(dim)
        coords = dict(self.coords)
        coords[dim] = coords[dim][idx]
        return DataArray(np.take(self.values, idx, axis=ax), self.dims, coords, self.name)

    # -- interpolation ---------------------------------------------------------------------------
    def interp(self, coords=None, method='linear', kwargs=None, **coords_kw):
        points = dict(coords or {}, **coords_kw)
        out = self
        for dim, new in points.items():                      # orthogonal: one dimension at a time
            new = np.atleast_1d(_raw(new)).astype(float)
            ax = out.dims.index(dim)
            f = interp1d(out.coords[dim].astype(float), out.values, kind=method, axis=ax,
                         bounds_error=False, assume_sorted=False, copy=False, **(kwargs or {}))
            c = dict(out.coords)
            c[dim] = new
            out = DataArray(f(new), out.dims, c, out.name)
        return out

    def integrate(self, dim):
        ax = self.dims.index(dim)
        x = self.coords[dim].astype(float)
        trapz = getattr(np, 'trapezoid', None) or np.trapz
        coords = {k: v for k, v in self.coords.items() if k != dim}
        return DataArray(trapz(self.values, x, axis=ax), self.dims[:ax] + self.dims[ax + 1:], coords, self.name)

    def expand_dims(self, mapping):
        (dim, val), = mapping.items()
        val = np.asarray([float(_raw(v)) for v in val])
        coords = dict(self.coords)
        coords[dim] = val
        return DataArray(self.values[None, ...], (dim,) + self.dims, coords, self.name)

    def drop_duplicates(self, dim):
        ax = self.dims.index(dim)
        _, first = np.unique(self.coords[dim], return_index=True)
        keep = np.sort(first)
        return self.isel_dim(dim, keep)


class Dataset:
    def __init__(self, data_vars=None, coords=None):
        self.coords = {}
        for k, v in (coords or {}).items():
            self.coords[k] = np.asarray(v[1] if isinstance(v, tuple) else v)
        self.data_vars = {}
        for k, (dims, arr) in (data_vars or {}).items():
            self.data_vars[k] = DataArray(arr, dims, {d: self.coords[d] for d in dims}, name=k)

    def __getattr__(self, name):
        d = self.__dict__
        if name in d.get('data_vars', {}):
            return d['data_vars'][name]
        if name in d.get('coords', {}):
            c = d['coords'][name]
            return DataArray(c, (name,), {name: c}, name=name)
        raise AttributeError(name)

    def interp(self, coords=None, method='linear', kwargs=None, **kw):
        out = Dataset()
        out.data_vars = {k: v.interp(coords, method=method, kwargs=kwargs, **kw) for k, v in self.data_vars.items()}
        first = next(iter(out.data_vars.values()))
        out.coords = dict(first.coords)
        return out

    def groupby_bins(self, name, bins):
        return _GroupByBins(self, name, np.asarray(getattr(bins, 'value', bins), dtype=float))

    def to_netcdf(self, path, encoding=None):
        Dataset.written.append((path, self, encoding))
    written = []

    def save_npz(self, path):
        """Test helper: what the reference would read back with xr.open_dataset."""
        (name, da), = self.data_vars.items()
        with open(path, 'wb') as fh:
            np.savez(fh, name=name, dims=np.array(da.dims), values=da.values,
                     **{'coord_' + d: self.coords[d] for d in da.dims})


class _GroupByBins:
    def __init__(self, ds, name, bins):
        self.ds, self.name, self.bins = ds, name, bins

    def map(self, func, **kw):
        codes = np.asarray(pd.cut(self.ds.coords[self.name], self.bins).codes)
        parts = []
        for b in range(len(self.bins) - 1):                   # bin order; empty bins are skipped
            idx = np.flatnonzero(codes == b)
            if idx.size == 0:
                continue
            g = Dataset()
            g.coords = dict(self.ds.coords)
            g.coords[self.name] = self.ds.coords[self.name][idx]
            g.data_vars = {k: v.isel_dim(self.name, idx) for k, v in self.ds.data_vars.items()}
            parts.append(func(g, **kw))
        return concat(parts, self.name)


def concat(arrays, dim):
    first = arrays[0]
    if dim in first.dims:
        ax = first.dims.index(dim)
        coords = dict(first.coords)
        coords[dim] = np.concatenate([a.coords[dim] for a in arrays])
        return DataArray(np.concatenate([a.values for a in arrays], axis=ax), first.dims, coords, first.name)
    return DataArray(np.stack([a.values for a in arrays], axis=0), (dim,) + first.dims, dict(first.coords))


def open_dataset(path, **kw):
    with np.load(path, allow_pickle=False) as z:
        dims = tuple(str(d) for d in z['dims'])
        return Dataset(data_vars={str(z['name']): (dims, z['values'])},
                       coords={d: z['coord_' + d] for d in dims})


def apply_ufunc(func, *args, input_core_dims=None, output_core_dims=None, output_dtypes=None,
                dask_gufunc_kwargs=None, kwargs=None, dask=None):
    raws = []
    lead = None
    for a, core in zip(args, input_core_dims):
        order = [a.dims.index(d) for d in a.dims if d not in core] + [a.dims.index(d) for d in core]
        raws.append(np.transpose(a.values, order))
        if lead is None and len(a.dims) > len(core):
            lead = [d for d in a.dims if d not in core]
            lead_coords = {d: a.coords[d] for d in lead if d in a.coords}
    out = func(*raws, **(kwargs or {}))
    lead = lead or []
    return DataArray(out, tuple(lead) + tuple(output_core_dims[0]), dict(lead_coords if lead else {}))
