funcs_common = ['first', 'last', 'len', 'mean', 'var', 'std', 'allnan', 'anynan', 'max', 'min',
                'argmax', 'argmin', 'sumofsquares', 'cumsum', 'cumprod', 'cummax', 'cummin',
                'sum', 'prod', 'all', 'any']
funcs_no_separate_nan = frozenset(['sort', 'rsort', 'array', 'allnan', 'anynan'])
