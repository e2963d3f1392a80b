"""
Generates tests/golden/*.json.

oracle_default_grid.json — outputs of the oracle on the reference's own test
  grid (frei/tests/test_core.py:19-46: default Grid, T_ref = 2400 K,
  load_example_opacity(scale_factor=1), emission_spectrum(n_timesteps=1)) with
  an H2O VMR of 3e-4, recorded at the commit where the oracle was validated
  against the reference's three known-answer values.  It guards the oracle (and
  through it the GPU path) against accidental edits.

reference_run_*.json — outputs of the reference's OWN source files
  (frei/twostream.py, frei/opacity.py, frei/core.py, frei/tp.py,
  frei/chemistry.py) executed under the dependency stubs of
  tests/golden/refstubs/ (astropy.units / xarray / specutils / tqdm shims,
  because those packages are not installable here).  Only regenerated when
  /root/reference exists; see tests/golden/run_reference.py.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import frei_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    pl = O.hot_jupiter()
    P = O.pressure_grid(30, np.log10(1e-6), np.log10(200))
    T = O.temperature_grid(P, 2400.0, 0.1, 0.1)
    lam, _, _ = O.wavelength_grid(0.5, 10, 500)
    tabs = O.load_example_opacity(P, T, lam, scale_factor=1)
    vmr = 3e-4
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'], vmr=vmr)
    spec, Tf, hist, dtaus, n_it = O.emission_spectrum(tabs, T, P, lam, pl, lambda a, b: mmr,
                                                      n_timesteps=1)
    k, s = O.kappa(tabs, T[0], P[0], lam, mmr, pl['m_bar'])
    idx = list(range(0, 500, 10))
    out = dict(vmr=vmr, lam_index=idx, spectrum=spec[idx].tolist(), final_temps=Tf.tolist(),
               dtaus=dtaus[:, idx].tolist(), kappa_layer0=k[idx].tolist(), sigma=s[idx].tolist(),
               peak_index=int(spec.argmax()), peak_flux=float(spec.max()))
    with open(os.path.join(HERE, 'oracle_default_grid.json'), 'w') as fh:
        json.dump(out, fh)
    print('wrote oracle_default_grid.json: peak', lam[spec.argmax()], spec.max())


if __name__ == '__main__':
    main()
