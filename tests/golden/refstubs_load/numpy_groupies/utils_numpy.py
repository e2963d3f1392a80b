"""numpy_groupies.utils_numpy: aliasing table and the input validation with the axis form the
reference uses (1-D group_idx, n-D array, axis=-1: labels are offset per leading row so that the
flattened problem never mixes rows — numpy_groupies' offset_labels)."""
import numpy as np

from . import utils

_alias_numpy = {np.sum: 'sum', np.prod: 'prod', np.mean: 'mean', np.max: 'max', np.min: 'min'}


def get_aliasing(*extra):
    alias = {f: f for f in utils.funcs_common}
    alias.update({'nan' + f: 'nan' + f for f in utils.funcs_common if f not in utils.funcs_no_separate_nan})
    for d in extra:
        alias.update(d)
        alias.update({v: v for v in d.values()})
        alias.update({'nan' + v: 'nan' + v for v in d.values()})
    return alias


def input_validation(group_idx, a, size=None, order='C', axis=None, ravel_group_idx=True,
                     check_bounds=True, func=None):
    group_idx = np.asanyarray(group_idx)
    a = np.asanyarray(a)
    if not np.issubdtype(group_idx.dtype, np.integer):
        raise TypeError('group_idx must be of integer type')
    if axis is None:
        raise NotImplementedError('stub: only the axis form is used by frei')
    axis = a.ndim + axis if axis < 0 else axis
    if group_idx.ndim != 1 or a.shape[axis] != group_idx.size or axis != a.ndim - 1:
        raise NotImplementedError('stub: 1-D group_idx along the last axis only')
    if size is None:
        size = int(group_idx.max()) + 1
    lead = a.shape[:-1]
    rows = int(np.prod(lead)) if lead else 1
    labels = np.broadcast_to(group_idx, a.shape) + (np.arange(rows, dtype=int) * size).reshape(lead + (1,))
    out_shape = lead + (size,)
    ndim_idx = len(out_shape)
    return labels.ravel(), a.ravel(), rows * size, ndim_idx, out_shape, None
