"""Independent double-precision construction of the real spherical-harmonics basis the reference hard-codes
(shencoder/src/shencoder.cu:50-121).  TEST INFRASTRUCTURE ONLY.

The reference (via tiny-cuda-nn) lists the degree<=8 basis as 64 explicit polynomials.  Here the same functions are
built from the textbook definition -- associated Legendre functions and cos/sin(m*phi) with the Condon-Shortley
phase -- so a transcription error in either the oracle or the CUDA template shows up as a disagreement.
Valid on UNIT vectors (the polynomial forms use x^2+y^2+z^2 = 1), which is what view directions are.

    index = l*l + l + m,  m = -l..l
    Y_l^0  = K_l^0 P_l(z)
    Y_l^m  = (-1)^m sqrt(2) K_l^m cos(m phi) P_l^m(z)       m > 0   (P without the Condon-Shortley phase)
    Y_l^-m = (-1)^m sqrt(2) K_l^m sin(m phi) P_l^m(z)
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import lpmv


def real_sh(dirs: np.ndarray, degree: int) -> np.ndarray:
    d = np.asarray(dirs, dtype=np.float64).reshape(-1, 3)
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    phi = np.arctan2(y, x)
    out = np.zeros((d.shape[0], degree * degree), np.float64)
    for l in range(degree):
        for m in range(0, l + 1):
            K = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m))
            P = (-1) ** m * lpmv(m, l, z)  # strip scipy's Condon-Shortley phase
            if m == 0:
                out[:, l * l + l] = K * P
            else:
                s = (-1) ** m * math.sqrt(2.0) * K
                out[:, l * l + l + m] = s * np.cos(m * phi) * P
                out[:, l * l + l - m] = s * np.sin(m * phi) * P
    return out
