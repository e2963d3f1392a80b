"""
Chemistry entry points — same interface as ``frei/chemistry.py``.

``chemistry()`` produces the mass-mixing ratios ``mmr[layer, species]`` that
weight the opacity tables on the hot path (frei/opacity.py:246-263).  The
equilibrium solver itself is the third-party pyfastchem package; when it is
absent the reference substitutes a mock with a constant volume-mixing ratio of
1.5e-3 for every species (frei/chemistry.py:143-153, 207-246) and so does this
module.  This is host code: its output is an *input array* of the GPU path.
"""
import os
import re

import numpy as np

from . import units as U

__all__ = ['chemistry', 'iso_to_species', 'iso_to_mass', 'species_name_to_fastchem_name',
           'species_name_to_common_isotopologue_name']

MOCK_VMR = 1.5e-3          # frei/chemistry.py:243
FASTCHEM_UNKNOWN = 9999999

# standard atomic weights (u) for when ``periodictable`` is not installed
_ATOMIC_WEIGHT = {
    'H': 1.00794, 'He': 4.002602, 'Li': 6.941, 'Be': 9.012182, 'B': 10.811, 'C': 12.0107,
    'N': 14.0067, 'O': 15.9994, 'F': 18.9984032, 'Ne': 20.1797, 'Na': 22.98976928,
    'Mg': 24.305, 'Al': 26.9815386, 'Si': 28.0855, 'P': 30.973762, 'S': 32.065,
    'Cl': 35.453, 'Ar': 39.948, 'K': 39.0983, 'Ca': 40.078, 'Sc': 44.955912, 'Ti': 47.867,
    'V': 50.9415, 'Cr': 51.9961, 'Mn': 54.938045, 'Fe': 55.845, 'Co': 58.933195,
    'Ni': 58.6934, 'Cu': 63.546, 'Zn': 65.409, 'Ga': 69.723, 'Ge': 72.64, 'As': 74.9216,
    'Se': 78.96, 'Br': 79.904, 'Kr': 83.798, 'Rb': 85.4678, 'Sr': 87.62, 'Y': 88.90585,
    'Zr': 91.224, 'Nb': 92.90638, 'Mo': 95.94, 'Ru': 101.07, 'Rh': 102.9055, 'Pd': 106.42,
    'Ag': 107.8682, 'Cd': 112.411, 'In': 114.818, 'Sn': 118.71, 'Sb': 121.76, 'Te': 127.6,
    'I': 126.90447, 'Xe': 131.293, 'Cs': 132.9054519, 'Ba': 137.327, 'La': 138.90547,
    'Ce': 140.116, 'W': 183.84, 'Pt': 195.084, 'Au': 196.966569, 'Hg': 200.59, 'Pb': 207.2,
}


def _atomic_weight(symbol):
    try:
        from periodictable import elements
        return getattr(elements, symbol).mass
    except ImportError:
        return _ATOMIC_WEIGHT[symbol]


def _formula_tokens(name):
    """'ClAlF2' -> [('Cl', 1), ('Al', 1), ('F', 2)]."""
    return [(el, int(n) if n else 1) for el, n in re.findall(r'([A-Z][a-z]?)(\d*)', name)]


def iso_to_species(isotopologue):
    """'1H2-16O' -> 'H2O', '48Ti-16O' -> 'TiO', 'Na' -> 'Na' (frei/chemistry.py:13-21)."""
    out = ''.join(m.group(1) + m.group(2)
                  for part in isotopologue.split('-')
                  for m in [re.match(r'\d*([A-Za-z]+)(\d*)$', part)] if m)
    return out if out else isotopologue


def iso_to_mass(isotopologue):
    """
    Mass in u from the isotope numbers: '1H2-16O' -> 18, '48Ti-16O' -> 64; a bare
    element symbol falls back to its standard atomic weight (frei/chemistry.py:24-37).
    """
    mass = 0.0
    for part in isotopologue.split('-'):
        m = re.match(r'(\d+)[A-Za-z]+(\d*)$', part)
        if m:
            mass += float(m.group(1)) * (float(m.group(2)) if m.group(2) else 1.0)
    if mass == 0:
        mass = _atomic_weight(isotopologue)
    return mass * U._u.u if U.HAVE_ASTROPY else mass


def _mass_in_u(isotopologue):
    m = iso_to_mass(isotopologue)
    return float(m.to(U._u.u).value) if U.is_quantity(m) else float(m)


def species_name_to_fastchem_name(k, return_mass=False):
    """
    'H2O' -> 'H2O1', 'ClAlF2' -> 'Al1Cl1F2' (alphabetical element order with
    explicit counts); single atoms keep their bare symbol (frei/chemistry.py:40-76).
    """
    toks = _formula_tokens(k)
    name = ''.join(f'{el}{n}' for el, n in sorted(toks, key=lambda t: t[0]))
    if len(toks) == 1 and toks[0][1] == 1:
        name = toks[0][0]
    if return_mass:
        return name, sum(_atomic_weight(el) * n for el, n in toks)
    return name


def species_name_to_common_isotopologue_name(k):
    """'H2O' -> '1H2-16O', 'AlClF2' -> '27Al-35Cl-19F2', 'Na' -> 'Na' (frei/chemistry.py:79-111)."""
    toks = _formula_tokens(k)
    if len(toks) <= 1:
        return toks[0][0] if toks else k
    return '-'.join(f'{round(_atomic_weight(el))}{el}{n if n > 1 else ""}' for el, n in toks)


def chemistry(temperatures, pressures, species, return_vmr=False, m_bar=2.4 * U.m_p):
    """
    Mass-mixing ratio of each species at every (T, P): dict isotopologue -> array.
    Same signature and return convention as frei/chemistry.py:114-205.
    """
    T = np.atleast_1d(U.value(temperatures, 'K'))
    P = np.atleast_1d(U.value(pressures, 'bar'))
    m_bar_g = float(U.value(m_bar, 'g'))
    species = list(species)
    pyfastchem = U.optional_module('pyfastchem')

    vmrs = {}
    if pyfastchem is None:
        # Mock: number density = 1.5e-3 * P / (k_B T) for every species (chemistry.py:232-246)
        n_gas = (P * 1e6) / (U.k_B * T)
        for iso in species:
            vmrs[iso] = (MOCK_VMR * n_gas) / n_gas
    else:                                                     # pragma: no cover
        data = os.path.join(os.path.dirname(pyfastchem.__file__), 'input')
        here = os.environ.get('FREI_FASTCHEM_DATA', data)
        fc = pyfastchem.FastChem(os.path.join(here, 'element_abundances_solar.dat'),
                                 os.path.join(here, 'logK.dat'), 0)
        inp, out = pyfastchem.FastChemInput(), pyfastchem.FastChemOutput()
        inp.temperature = T[::-1]
        inp.pressure = P[::-1]
        fc.calcDensities(inp, out)
        n = np.array(out.number_densities)
        n_gas = (P[::-1] * 1e6) / (U.k_B * T[::-1])
        for iso in species:
            idx = fc.getSpeciesIndex(species_name_to_fastchem_name(iso_to_species(iso)))
            if idx != pyfastchem.FASTCHEM_UNKNOWN_SPECIES:
                vmrs[iso] = (n[:, idx] / n_gas)[::-1]
            else:
                print("Species", iso_to_species(iso), "not found in FastChem")
    mmrs = {iso: v * (_mass_in_u(iso) * U.amu / m_bar_g) for iso, v in vmrs.items()}
    if return_vmr:
        return mmrs, vmrs
    return mmrs
