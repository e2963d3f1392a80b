"""Stand-in for specutils.Spectrum1D (flux + spectral axis), see astropy stub."""


class Spectrum1D:
    def __init__(self, flux=None, spectral_axis=None):
        self.flux, self.spectral_axis = flux, spectral_axis

    @property
    def wavelength(self):
        return self.spectral_axis
