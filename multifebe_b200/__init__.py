"""multifebe_b200 -- B200-native hot path of MultiFEBE's time-harmonic 3D elastodynamic BEM
(influence-matrix assembly + dense complex LU) behind a C ABI (include/mfb.h)."""
__version__ = "0.1.0"
