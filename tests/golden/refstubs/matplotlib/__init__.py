"""Stand-in for matplotlib used ONLY to execute the reference's own frei/plot.py in
tests/golden/run_reference.py (matplotlib is not installable here): every drawing call is
accepted and ignored, except Axes.pcolormesh, which hands the plotted array (the normalised
contribution function, frei/plot.py:83) to whoever registered `capture` and stops the routine."""


class Captured(Exception):
    def __init__(self, args):
        super().__init__('pcolormesh captured')
        self.payload = args


class _Anything:
    """Accepts any attribute access, call, indexing or context-manager use."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __getitem__(self, key):
        return _Anything()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __iter__(self):
        return iter(())


class Axes(_Anything):
    def pcolormesh(self, *args, **kwargs):
        raise Captured(args)


class Figure(_Anything):
    def add_subplot(self, *a, **k):
        return Axes()
