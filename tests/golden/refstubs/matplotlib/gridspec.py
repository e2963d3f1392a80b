from . import _Anything


class GridSpec(_Anything):
    def __init__(self, *a, **k):
        pass
