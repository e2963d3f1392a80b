"""
Stand-in for the handful of xarray features frei/opacity.py uses on the hot path.
DataArray.interp mirrors what xarray does for a point-wise (shared 'z' dim) linear
interpolation: coordinates sorted ascending, scipy.interpolate.interpn with
bounds_error=False and the given fill_value.
"""
import numpy as np
from scipy.interpolate import interpn


class DataArray:
    __array_ufunc__ = None            # make ndarray.__mul__ defer to __rmul__

    def __init__(self, data, dims=None, coords=None):
        self.values = np.asarray(data, dtype=float)
        self.dims = (dims,) if isinstance(dims, str) else tuple(dims or ())
        self.coords = {k: np.asarray(v, dtype=float) for k, v in (coords or {}).items()}

    def __getattr__(self, name):
        coords = self.__dict__.get('coords', {})
        if name in coords:
            return coords[name]
        raise AttributeError(name)

    def drop_duplicates(self, dim):
        ax = self.dims.index(dim)
        _, first = np.unique(self.coords[dim], return_index=True)
        keep = np.sort(first)
        coords = dict(self.coords)
        coords[dim] = coords[dim][keep]
        return DataArray(np.take(self.values, keep, axis=ax), self.dims, coords)

    def interp(self, method='linear', kwargs=None, **points):
        assert method == 'linear' and set(points) == {'pressure', 'temperature'}, 'stub: 2-D only'
        fill = (kwargs or {}).get('fill_value', np.nan)
        order = [self.dims.index(d) for d in ('pressure', 'temperature', 'wavelength')]
        vals = np.transpose(self.values, order)
        P, T = self.coords['pressure'], self.coords['temperature']
        ip, it = np.argsort(P, kind='stable'), np.argsort(T, kind='stable')
        vals = vals[ip][:, it]
        xi = np.stack([points['pressure'].values, points['temperature'].values], axis=-1)
        out = interpn((P[ip], T[it]), vals, xi, method='linear', bounds_error=False, fill_value=fill)
        return DataArray(out, ('z', 'wavelength'), dict(wavelength=self.coords['wavelength']))

    def __rmul__(self, other):
        return DataArray(np.asarray(other) * self.values, self.dims, self.coords)
    __mul__ = __rmul__

    def sum(self, dim):
        ax = self.dims.index(dim)
        return DataArray(self.values.sum(axis=ax), self.dims[:ax] + self.dims[ax + 1:], self.coords)


def concat(arrays, dim):
    return DataArray(np.stack([a.values for a in arrays], axis=0), (dim,) + arrays[0].dims,
                     arrays[0].coords)
