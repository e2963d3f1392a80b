"""
A small re-implementation of the parts of astropy.units the reference's hot path uses.
Semantics follow astropy: a Quantity is an ndarray holding the value in its own unit; units
compose symbolically as (scale to SI, exponents of m, kg, s, K); conversions happen on .to(),
.decompose(), addition/comparison of different units and item assignment.
"""
import functools

import numpy as np

__all__ = ['Unit', 'Quantity', 'quantity_input', 'spectral_density', 'spectral', 'dimensionless_unscaled']


def _dims_eq(a, b):
    return all(abs(x - y) < 1e-9 for x, y in zip(a, b))


class Unit:
    __array_ufunc__ = None            # ndarray (op) Unit defers to Unit.__r<op>__

    def __init__(self, scale, dims, name=''):
        self.scale, self.dims, self.name = float(scale), tuple(float(d) for d in dims), name

    # unit algebra
    def __mul__(self, o):
        if isinstance(o, Unit):
            return Unit(self.scale * o.scale, [a + b for a, b in zip(self.dims, o.dims)])
        return Quantity(o, self)
    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Unit):
            return Unit(self.scale / o.scale, [a - b for a, b in zip(self.dims, o.dims)])
        return Quantity(1.0 / np.asarray(o, dtype=float), self)

    def __rtruediv__(self, o):
        return Quantity(o, self ** -1)

    def __pow__(self, p):
        return Unit(self.scale ** p, [d * p for d in self.dims])

    def is_dimensionless(self):
        return _dims_eq(self.dims, (0, 0, 0, 0))

    def factor_to(self, other):
        if not _dims_eq(self.dims, other.dims):
            raise ValueError(f'unit mismatch {self.dims} vs {other.dims}')
        return self.scale / other.scale

    def __repr__(self):
        return self.name or f'Unit({self.scale:g}, {self.dims})'


def _u(other):
    return other.unit if isinstance(other, Quantity) else dimensionless_unscaled


class Quantity(np.ndarray):
    __array_priority__ = 10000

    def __new__(cls, value, unit=None, dtype=None, copy=True):
        if isinstance(value, Quantity):
            if unit is None:
                unit = value.unit
            arr = np.array(value.view(np.ndarray), dtype=float) * value.unit.factor_to(unit)
        elif isinstance(value, (list, tuple)) and len(value) and isinstance(value[0], Quantity):
            unit0 = value[0].unit if unit is None else unit
            arr = np.array([np.asarray(v.to(unit0).value) for v in value], dtype=float)
            unit = unit0
        else:
            arr = np.array(value, dtype=float)
        obj = arr.view(cls)
        obj.unit = dimensionless_unscaled if unit is None else unit
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, 'unit', None) or dimensionless_unscaled

    # basics -----------------------------------------------------------------
    @property
    def value(self):
        v = self.view(np.ndarray)
        return v if v.ndim else v[()]

    @property
    def isscalar(self):
        return self.ndim == 0

    def to(self, unit, equivalencies=None):
        try:
            f = self.unit.factor_to(unit)
        except ValueError:
            if equivalencies is None:
                raise
            if isinstance(equivalencies, _Spectral):          # wavelength <-> wavenumber: 1 / lambda
                return Quantity(1.0 / self.view(np.ndarray), self.unit ** -1).to(unit)
            return (self * equivalencies.lam).to(unit)        # F_lambda -> lambda F_lambda
        return Quantity(self.view(np.ndarray) * f, unit)

    def decompose(self):
        return Quantity(self.view(np.ndarray) * self.unit.scale, Unit(1.0, self.unit.dims))

    def copy(self):
        return Quantity(self.view(np.ndarray).copy(), self.unit)

    def flatten(self):
        return Quantity(self.view(np.ndarray).flatten(), self.unit)

    def __getitem__(self, key):
        return Quantity(self.view(np.ndarray)[key], self.unit)

    def __setitem__(self, key, val):
        if isinstance(val, Quantity):
            val = val.view(np.ndarray) * val.unit.factor_to(self.unit)
        self.view(np.ndarray)[key] = val

    def __iter__(self):
        v = self.view(np.ndarray)
        return (Quantity(x, self.unit) for x in v)

    @property
    def T(self):
        return Quantity(self.view(np.ndarray).T, self.unit)

    def __format__(self, spec):
        return format(float(self.value), spec) if self.ndim == 0 else str(self)

    def __repr__(self):
        return f'<Quantity {self.view(np.ndarray)!r} {self.unit!r}>'
    __str__ = __repr__

    def __bool__(self):
        return bool(self.view(np.ndarray))

    def __float__(self):
        return float(self.to(dimensionless_unscaled).view(np.ndarray))

    def __pow__(self, p):
        return Quantity(self.view(np.ndarray) ** p, self.unit ** p)

    def __mul__(self, o):
        if isinstance(o, Unit):
            return Quantity(self.view(np.ndarray), self.unit * o)
        return np.multiply(self, o)
    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Unit):
            return Quantity(self.view(np.ndarray), self.unit / o)
        return np.true_divide(self, o)

    # numpy protocol ----------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        raw = [x.view(np.ndarray) if isinstance(x, Quantity) else x for x in inputs]
        name = ufunc.__name__
        if 'out' in kwargs:                                   # in-place ops (q += x)
            target = kwargs.pop('out')[0]
            res = self.__array_ufunc__(ufunc, method, *inputs, **kwargs)
            if name in ('multiply', 'true_divide', 'divide') and isinstance(res, Quantity) \
                    and isinstance(target, Quantity):
                # astropy: q *= x and q /= x change the unit of q in place
                target.view(np.ndarray)[...] = res.view(np.ndarray)
                target.unit = res.unit
                return target
            target[...] = res
            return target
        if method == 'reduce':
            out = getattr(ufunc, method)(raw[0], **kwargs)
            return Quantity(out, inputs[0].unit) if name in ('add', 'maximum', 'minimum') else out
        if method != '__call__':
            return NotImplemented
        if name in ('multiply',):
            return Quantity(ufunc(*raw, **kwargs), _u(inputs[0]) * _u(inputs[1]))
        if name in ('true_divide', 'divide'):
            return Quantity(ufunc(*raw, **kwargs), _u(inputs[0]) / _u(inputs[1]))
        if name in ('add', 'subtract', 'minimum', 'maximum', 'greater', 'less', 'greater_equal',
                    'less_equal', 'equal', 'not_equal'):
            ua = _u(inputs[0])
            b = raw[1]
            ub = _u(inputs[1])
            if not isinstance(inputs[0], Quantity):          # plain (op) Quantity
                ua = ub if not ub.is_dimensionless() and np.all(np.asarray(raw[0]) == 0) else ua
            if not isinstance(inputs[1], Quantity) and np.all(np.asarray(b) == 0):
                ub = ua                                       # comparisons with a bare 0
            b = np.asarray(b, dtype=float) * ub.factor_to(ua)
            out = ufunc(raw[0], b, **kwargs)
            return Quantity(out, ua) if name in ('add', 'subtract', 'minimum', 'maximum') else out
        if name in ('negative', 'absolute', 'fabs', 'positive'):
            return Quantity(ufunc(*raw, **kwargs), inputs[0].unit)
        if name == 'power':
            return Quantity(ufunc(*raw, **kwargs), inputs[0].unit ** float(raw[1]))
        if name == 'sqrt':
            return Quantity(ufunc(*raw, **kwargs), inputs[0].unit ** 0.5)
        if name in ('exp', 'expm1', 'log', 'log10', 'sign', 'isfinite', 'isnan'):
            x = inputs[0]
            if name in ('sign', 'isfinite', 'isnan'):
                return ufunc(raw[0], **kwargs)
            return Quantity(ufunc(x.view(np.ndarray) * x.unit.factor_to(dimensionless_unscaled),
                                  **kwargs), dimensionless_unscaled)
        raise NotImplementedError(f'ufunc {name} on Quantity')

    def __array_function__(self, func, types, args, kwargs):
        name = func.__name__
        if name in ('trapz', 'trapezoid'):
            y = args[0]
            x = args[1] if len(args) > 1 else kwargs.pop('x')
            kw = {k: v for k, v in kwargs.items() if k in ('axis',)}
            raw = lambda a: a.view(np.ndarray) if isinstance(a, Quantity) else np.asarray(a)
            out = np.trapezoid(raw(y), raw(x), **kw)
            return Quantity(out, _u(y) * _u(x))
        if name in ('hstack', 'concatenate', 'vstack', 'stack'):
            seq = list(args[0])
            unit = seq[0].unit
            out = func([q.to(unit).view(np.ndarray) for q in seq], *args[1:], **kwargs)
            return Quantity(out, unit)
        if name == 'interp':
            x, xp, fp = args[:3]
            conv = lambda a: a.view(np.ndarray) if isinstance(a, Quantity) else a
            out = func(conv(x), conv(xp), conv(fp), **kwargs)
            return Quantity(out, fp.unit) if isinstance(fp, Quantity) else out
        if name in ('all', 'any', 'count_nonzero', 'unique', 'diff', 'argmax', 'shape', 'ndim', 'size'):
            conv = [a.view(np.ndarray) if isinstance(a, Quantity) else a for a in args]
            return func(*conv, **kwargs)
        if name in ('zeros_like', 'ones_like', 'copy', 'max', 'amax', 'min', 'amin', 'mean', 'sum'):
            out = func(args[0].view(np.ndarray), *args[1:], **kwargs)
            return Quantity(out, args[0].unit)
        if name == 'where':
            conv = [a.view(np.ndarray) if isinstance(a, Quantity) else a for a in args]
            unit = next((a.unit for a in args[1:] if isinstance(a, Quantity)), None)
            out = func(*conv, **kwargs)
            return Quantity(out, unit) if unit is not None else out
        conv = [a.view(np.ndarray) if isinstance(a, Quantity) else a for a in args]
        return func(*conv, **kwargs)

    def max(self, *a, **k):
        return Quantity(self.view(np.ndarray).max(*a, **k), self.unit)

    def min(self, *a, **k):
        return Quantity(self.view(np.ndarray).min(*a, **k), self.unit)

    def mean(self, *a, **k):
        return Quantity(self.view(np.ndarray).mean(*a, **k), self.unit)

    def argmax(self, *a, **k):
        return self.view(np.ndarray).argmax(*a, **k)


class _SpectralDensity:
    def __init__(self, lam):
        self.lam = lam


def spectral_density(lam):
    return _SpectralDensity(lam)


class _Spectral:
    pass


def spectral():
    return _Spectral()


def quantity_input(*a, **kw):
    def deco(fn):
        return fn
    return deco if not (len(a) == 1 and callable(a[0])) else a[0]


# -- units (SI scale, exponents of m, kg, s, K) ----------------------------------------------
dimensionless_unscaled = Unit(1.0, (0, 0, 0, 0), '')
m = Unit(1.0, (1, 0, 0, 0), 'm')
cm = Unit(1e-2, (1, 0, 0, 0), 'cm')
km = Unit(1e3, (1, 0, 0, 0), 'km')
um = Unit(1e-6, (1, 0, 0, 0), 'um')
micron = um
kg = Unit(1.0, (0, 1, 0, 0), 'kg')
g = Unit(1e-3, (0, 1, 0, 0), 'g')
s = Unit(1.0, (0, 0, 1, 0), 's')
K = Unit(1.0, (0, 0, 0, 1), 'K')
J = Unit(1.0, (2, 1, -2, 0), 'J')
erg = Unit(1e-7, (2, 1, -2, 0), 'erg')
W = Unit(1.0, (2, 1, -3, 0), 'W')
Pa = Unit(1.0, (-1, 1, -2, 0), 'Pa')
bar = Unit(1e5, (-1, 1, -2, 0), 'bar')
u = Unit(1.66053906660e-27, (0, 1, 0, 0), 'u')
AU = Unit(1.495978707e11, (1, 0, 0, 0), 'AU')
R_sun = Unit(6.957e8, (1, 0, 0, 0), 'R_sun')
R_jup = Unit(7.1492e7, (1, 0, 0, 0), 'R_jup')
# astropy defines M_jup = GM_jup / G
M_jup = Unit(1.2668653e17 / 6.6743e-11, (0, 1, 0, 0), 'M_jup')
