"""Minimal astropy stand-in (units + constants) used ONLY to execute the reference's own source
files in tests/golden/run_reference.py: astropy is not installable in the build container."""
