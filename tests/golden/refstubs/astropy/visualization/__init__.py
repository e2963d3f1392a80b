"""quantity_support(): a no-op context manager (the plotting calls are stubs)."""
import contextlib


@contextlib.contextmanager
def quantity_support():
    yield
