"""CODATA 2018 / IAU 2015 constants as astropy >= 4.0 ships them (SI)."""
from . import units as u

h = u.Quantity(6.62607015e-34, u.J * u.s)
c = u.Quantity(299792458.0, u.m / u.s)
k_B = u.Quantity(1.380649e-23, u.J / u.K)
m_p = u.Quantity(1.67262192369e-27, u.kg)
sigma_sb = u.Quantity(5.6703744191844314e-08, u.W / u.m ** 2 / u.K ** 4)
G = u.Quantity(6.6743e-11, u.m ** 3 / u.kg / u.s ** 2)
