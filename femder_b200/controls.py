"""Problem-setup holders with the attribute names ``FEM3D`` reads (input contract only).

The reference's own ``AirProperties`` / ``AlgControls`` / ``BC`` / ``Source`` / ``Receiver`` objects
(``femder/controlsair.py:6-24, 90-113``, ``BoundaryConditions.py:3-90``, ``sources.py``, ``receivers.py``) can be
passed to :class:`femder_b200.FEM3D` unchanged - it only reads ``c0, rho0``, ``freq, w``, ``mu, v, rhoc, cc``,
``coord, q``.  These minimal equivalents exist so the package is usable without the reference installed.
"""
from __future__ import annotations

import numpy as np


class AirProperties:
    def __init__(self, c0=343.0, rho0=1.21, temperature=20.0, humid=50.0, p_atm=101325.0):
        self.c0 = np.array(c0)
        self.rho0 = np.array(rho0)
        self.temperature = np.array(temperature, dtype=np.float32)
        self.hr = np.array(humid, dtype=np.float32)
        self.p_atm = np.array(p_atm, dtype=np.float32)


class AlgControls:
    def __init__(self, AP, freq_init=100.0, freq_end=10000.0, freq_step=10, freq_vec=()):
        freq_vec = np.array(freq_vec, dtype=np.float64)
        self.c0 = AP.c0
        if freq_vec.size == 0:
            self.freq_init = np.array(freq_init)
            self.freq_end = np.array(freq_end)
            self.freq_step = np.array(freq_step)
            self.freq = np.arange(self.freq_init, self.freq_end + self.freq_step, self.freq_step)
        else:
            self.freq_init = np.array(freq_vec[0])
            self.freq_end = np.array(freq_vec[-1])
            self.freq = freq_vec
        self.w = 2.0 * np.pi * self.freq
        self.k0 = self.w / self.c0
        self.fcenter = self.freq


class BC:
    """Admittance per surface group per frequency: ``mu[group] -> (n_freq,)`` (real or complex)."""

    def __init__(self, AC, AP):
        self.AP = AP
        self.AC = AC
        self.mu = {}
        self.v = {}
        self.rhoc = {}
        self.cc = {}

    def rigid(self, domain_index):
        self.mu[domain_index] = np.zeros_like(self.AC.freq)

    def admittance(self, domain_index, admittance):
        self.mu[domain_index] = np.ones_like(self.AC.freq) * admittance

    def normalized_admittance(self, domain_index, normalized_admittance):
        idx = domain_index if isinstance(domain_index, (list, tuple)) else [domain_index]
        for i in idx:
            self.mu[i] = np.ones_like(self.AC.freq) * normalized_admittance / (self.AP.c0 * self.AP.rho0)

    def velocity(self, domain_index, velocity):
        """Normal velocity on a surface group (``BoundaryConditions.py:98-116``): fills ``v`` only, never ``mu``."""
        self.v[domain_index] = np.ones_like(self.AC.freq) * velocity

    def fluid(self, domain_index, cc, rhoc):
        """Equivalent fluid in a VOLUME domain (``BoundaryConditions.py:140-160``): complex sound speed and density."""
        self.rhoc[domain_index] = (np.ones_like(self.AC.freq, dtype=complex) * rhoc).ravel()
        self.cc[domain_index] = (np.ones_like(self.AC.freq, dtype=complex) * cc).ravel()

    def impedance(self, domain_index, impedance):
        idx = domain_index if isinstance(domain_index, (list, tuple)) else [domain_index]
        for i in idx:
            self.mu[i] = np.ones_like(self.AC.freq) / np.asarray(impedance)


class Source:
    def __init__(self, wavetype="spherical", coord=(0.0, 0.0, 1.0), q=(1.0,)):
        self.coord = np.reshape(np.array(coord, dtype=np.float64), (-1, 3))
        self.q = np.array([q], dtype=np.float64).reshape(-1, 1)
        self.wavetype = wavetype


class Receiver:
    def __init__(self, coord=(1.0, 0.0, 0.0)):
        self.coord = np.reshape(np.array(coord, dtype=np.float64), (-1, 3))
