"""numpy_groupies.aggregate_numba: names frei/interp.py imports.  The aggregation classes other
than the reference's own Trapz are never called on frei's path; they only have to be constructible
(get_funcs, frei/interp.py:205-216, instantiates all of them)."""
import numpy as np

from .utils import funcs_no_separate_nan  # noqa: F401


class _Unused:
    def __init__(self, func=None, **kwargs):
        self.func = func

    def __call__(self, *a, **k):
        raise NotImplementedError(f'{self.func}: not part of the stub (frei only calls trapz)')


(Sum, Prod, Len, All, Any, Last, First, AllNan, AnyNan, Min, Max, ArgMin, ArgMax, Mean, Std, Var,
 CumSum, CumProd, CumMax, CumMin) = [type(n, (_Unused,), {}) for n in (
     'Sum', 'Prod', 'Len', 'All', 'Any', 'Last', 'First', 'AllNan', 'AnyNan', 'Min', 'Max', 'ArgMin',
     'ArgMax', 'Mean', 'Std', 'Var', 'CumSum', 'CumProd', 'CumMax', 'CumMin')]

_default_cache = {}


def isstr(s):
    return isinstance(s, str)


def check_dtype(dtype, func_str, a, n):
    """Result dtype: the given one, else the input's floating type (float64 for integers)."""
    if dtype is not None:
        return np.dtype(dtype)
    a_dtype = np.dtype(type(a)) if np.isscalar(a) else a.dtype
    return a_dtype if np.issubdtype(a_dtype, np.floating) else np.dtype(np.float64)


def check_fill_value(fill_value, dtype, func=None):
    try:
        return dtype.type(fill_value)
    except ValueError:
        raise ValueError(f'fill_value must be convertible into {dtype.type.__name__}')


def get_func(func, aliasing, implementations):
    try:
        func_str = aliasing[func]
    except (KeyError, TypeError):
        if callable(func):
            return func
    else:
        if func_str in implementations:
            return func_str
        raise NotImplementedError('No such function available')
    raise ValueError(f'func {func} is neither a valid function string nor a callable object')
