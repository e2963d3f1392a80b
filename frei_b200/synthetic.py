"""
Seeded synthetic workloads: the inputs fed identically to
the GPU path, the CPU oracle and the CPU baseline.  Pure input generation — no
part of the hot path lives here.

Opacity table of species s:  kappa_s(P, T, lambda) = base_s(lambda) * fac_s(P, T),
  base_s = scale_s * 5 * 10^(2.5 (so_s(lambda) - 0.4))  with so_s the Gaussian-band
  construction of the reference's fixture (frei/opacity.py:301-326) re-seeded
  with 42 + s, and fac_s = (T / 2000 K)^a_s (P / 1 bar)^b_s so that the (P, T)
  interpolation is non-trivial.  Axes: 12 pressure nodes 1e-7 .. 1e4 bar (one per
  decade), 24 temperature nodes 200 .. 9400 K (400 K steps).
"""
import numpy as np

from . import units as U

SPECIES = ['1H2-16O', '12C-16O', '12C-1H4', '12C-16O2', '14N-1H3', '48Ti-16O', '51V-16O',
           '1H-12C-14N']
SPECIES_MASS_U = [18.0, 28.0, 16.0, 44.0, 17.0, 64.0, 67.0, 27.0]

CONFIGS = {
    # name: (n_layers, n_lambda, n_species, T_ref)
    'C1': (50, 5_000, 3, 2400.0),
    'C2': (50, 200_000, 3, 2400.0),
    'C3': (100, 1_000_000, 8, 3200.0),
    'C4': (50, 20_000, 3, 2400.0),       # per atmosphere; batch of 4096
    'C5': (200, 2_000_000, 3, 2400.0),
}


def band_profile(lam_um, seed=42):
    """The ``so`` curve of the reference fixture for one seed (frei/opacity.py:295-324)."""
    rs = np.random.RandomState(seed)            # same stream as np.random.seed(seed)
    amps = rs.uniform(low=0.1, high=0.2, size=15)
    cens = rs.uniform(low=0.5, high=1, size=15)
    so = np.exp(-0.5 * (lam_um - 6) ** 2 / 2 ** 2) + 0.8 * np.exp(-0.5 * (lam_um - 0.3) ** 2 / 0.5 ** 2)
    for amp, wl in zip(amps, cens):
        so = so + amp * np.exp(-0.5 * (lam_um - wl) ** 2 / 0.005 ** 2)
    for amp, wl in zip([0.22, 0.2, 0.18], np.logspace(np.log10(1.4), np.log10(2.7), 3)):
        so = so + amp * np.exp(-0.5 * (lam_um - wl) ** 2 / 0.13 ** 2)
    return so


def hot_jupiter_cgs():
    """Planet.from_hot_jupiter() in CGS numbers (frei/core.py:92-106)."""
    return dict(a_rstar=float(0.03 * U.au / U.R_sun), m_bar=2.4 * U.m_p,
                g=U.GM_jup / U.R_jup ** 2, T_star=5800.0, alpha=1)


def make_workload(n_layers, n_lam, n_species, T_ref=2400.0, table_f32=False, seed=0):
    """
    Returns a dict with lam_um, P_bar, T_init, planet (CGS), species, mmr[L,S],
    axis_P[N_P], axis_T[N_T], base[S, n_lam], fac[S, N_P, N_T].
    ``table_f32``: round base*fac products to float32-representable values so an
    fp32 device table holds exactly what the fp64 oracle sees.
    """
    lam_um = np.logspace(np.log10(0.5), np.log10(10), n_lam)
    P = np.logspace(-6, np.log10(200), n_layers)[::-1].copy()
    T = T_ref * (P / 0.1) ** 0.1
    planet = hot_jupiter_cgs()
    species = SPECIES[:n_species]
    masses = np.array(SPECIES_MASS_U[:n_species])
    mmr = np.broadcast_to(1.5e-3 * (masses * U.amu / planet['m_bar']), (n_layers, n_species)).copy()
    axis_P = 10.0 ** np.arange(-7, 5, dtype=np.float64)            # 12 nodes
    axis_T = 200.0 + 400.0 * np.arange(24, dtype=np.float64)       # 24 nodes
    rs = np.random.RandomState(1234 + seed)
    a = rs.uniform(-1, 1, n_species)
    b = rs.uniform(0, 0.3, n_species)
    scale = rs.uniform(0.5, 2.0, n_species)
    base = np.stack([scale[s] * 5 * 10 ** (2.5 * (band_profile(lam_um, 42 + s) - 0.4))
                     for s in range(n_species)])
    fac = np.stack([(axis_T[None, :] / 2000.0) ** a[s] * (axis_P[:, None] / 1.0) ** b[s]
                    for s in range(n_species)])
    return dict(lam_um=lam_um, P_bar=P, T_init=T, planet=planet, species=species, mmr=mmr,
                axis_P=axis_P, axis_T=axis_T, base=base, fac=fac, table_f32=table_f32,
                L=n_layers, n_lam=n_lam, S=n_species)


def host_tables(w, lam_index=None):
    """Oracle-style tables {species: dict(P, T, values[N_P, N_T, n_lam])} (optionally a lambda subset)."""
    out = {}
    for s, name in enumerate(w['species']):
        base = w['base'][s] if lam_index is None else w['base'][s][lam_index]
        vals = w['fac'][s][:, :, None] * base[None, None, :]
        if w['table_f32']:
            vals = vals.astype(np.float32).astype(np.float64)
        out[name] = dict(P=w['axis_P'], T=w['axis_T'], values=vals)
    return out


def device_table(w, dtype, lam_range=None, device=None):
    """Build the same table directly in HBM (outer product on the device)."""
    import torch
    from .engine import DeviceTable, FREI_F32
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
    lo, hi = (0, w['n_lam']) if lam_range is None else lam_range
    base = torch.from_numpy(np.ascontiguousarray(w['base'][:, lo:hi])).to(dev)
    fac = torch.from_numpy(w['fac']).to(dev)
    tdt = torch.float32 if dtype == FREI_F32 else torch.float64
    S, N_P, N_T = fac.shape
    vals = torch.empty((S, N_P, N_T, hi - lo), dtype=tdt, device=dev)
    for s in range(S):
        for ip in range(N_P):
            prod = fac[s, ip][:, None] * base[s][None, :]
            vals[s, ip].copy_(prod.to(torch.float32) if (w['table_f32'] or tdt == torch.float32)
                              else prod)
    tab = DeviceTable.from_device_values(vals, w['axis_P'], w['axis_T'], species=w['species'])
    tab.lam_range = (lo, hi)
    return tab
