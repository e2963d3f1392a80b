from . import Figure, _Anything

cm = _Anything()


def figure(*a, **k):
    return Figure()


def colorbar(*a, **k):
    return _Anything()
