"""Problem-setup holders for the benches, tests and scripts of this repository - NOT part of the product package.

``femder_b200.FEM3D`` takes the reference's own ``AirProperties`` / ``AlgControls`` / ``BC`` / ``Source`` / ``Receiver`` objects
(``femder/controlsair.py:6-24, 90-113``, ``BoundaryConditions.py:3-160``, ``sources.py``, ``receivers.py``) unchanged: it only reads
``c0, rho0``, ``freq, w``, ``mu, v, rhoc, cc``, ``coord, q``.  On a machine without the reference installed (the GPU box) these
minimal stand-ins carry the same attributes.
"""
from __future__ import annotations

import numpy as np


class AirProperties:
    """``c0, rho0`` as 0-d arrays like the reference's (``controlsair.py:7-24``); the rest is carried for completeness."""

    def __init__(self, c0=343.0, rho0=1.21, temperature=20.0, humid=50.0, p_atm=101325.0):
        self.c0, self.rho0 = np.array(c0), np.array(rho0)
        self.temperature, self.hr, self.p_atm = (np.array(v, dtype=np.float32) for v in (temperature, humid, p_atm))


class AlgControls:
    """The frequency list: ``freq`` (Hz), ``w = 2 pi freq``, ``k0 = w / c0`` (``controlsair.py:90-113``) - either an explicit
    ``freq_vec`` or ``freq_init .. freq_end`` (inclusive) every ``freq_step``."""

    def __init__(self, AP, freq_init=100.0, freq_end=10000.0, freq_step=10, freq_vec=()):
        self.c0 = AP.c0
        explicit = np.array(freq_vec, dtype=np.float64)
        if explicit.size:
            self.freq = explicit
        else:
            self.freq_step = np.array(freq_step)
            self.freq = np.arange(freq_init, freq_end + freq_step, freq_step)
        self.freq_init, self.freq_end = np.array(self.freq[0] if explicit.size else freq_init), np.array(self.freq[-1] if explicit.size else freq_end)
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
