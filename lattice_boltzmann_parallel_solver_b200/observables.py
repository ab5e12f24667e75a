"""Host-side estimators of the reference's derived observables, wired to this package's outputs (SURVEY.md §8(f)
row 4). They restate, without matplotlib, what the reference computes inside its plotting helpers."""
import numpy as np

from .lattice_boltzmann_method import strouhal_number


def probe_speed(uxuy):
    """|u| at the probe from the (n, 2) device probe ring — what the drivers append every step
    (reference: src/experiments.py:703-704, np.linalg.norm of velocity[px, py])."""
    uxuy = np.asarray(uxuy, dtype=np.float64)
    return np.sqrt(uxuy[..., 0] ** 2 + uxuy[..., 1] ** 2)


def vortex_frequency(vel_at_p, cut=70000):
    """Arg-max FFT bin of the mean-free probe trace after the transient
    (reference: src/visualizations_utils.py:150-163)."""
    v = np.array(vel_at_p[cut:], dtype=np.float64)
    v -= np.mean(v)
    spectrum = np.fft.fft(v)
    freq = np.fft.fftfreq(len(v), 1)
    return float(np.abs(freq[np.argmax(np.abs(spectrum))]))


def strouhal_from_trace(vel_at_p, plate_size=40, inlet_velocity=0.1, cut=70000):
    """St = f d / u (reference: src/visualizations_utils.py:163-167, src/lattice_boltzmann_method.py:74-90)."""
    return float(strouhal_number(vortex_frequency(vel_at_p, cut), plate_size, inlet_velocity))


def decay_amplitude(field_min, field_max, offset=0.0):
    """The per-step amplitude sample of the shear-wave experiment (reference: src/experiments.py:181-193)."""
    return np.abs(field_min) - offset if np.abs(field_min) > np.abs(field_max) else np.abs(field_max) - offset


def viscosity_from_decay(amplitudes, epsilon, wavelength, peaks_only=False):
    """Fit eps * exp(-nu (2 pi / L)^2 t) to the amplitude series (reference: src/experiments.py:195-209;
    `peaks_only` is the density branch, which fits the local maxima)."""
    from scipy.optimize import curve_fit
    from scipy.signal import argrelextrema
    a = np.asarray(amplitudes, dtype=np.float64)
    t = np.arange(0, len(a))
    if peaks_only:
        idx = argrelextrema(a, np.greater)
        t, a = np.array(idx).squeeze(), a[idx]
    k2 = np.power(2 * np.pi / wavelength, 2)
    return float(curve_fit(lambda tt, v: epsilon * np.exp(-v * k2 * tt), t, a)[0][0])


def theoretical_viscosity(omega):
    """nu = (1/3)(1/omega - 1/2) (reference: src/experiments.py:210)."""
    return (1 / 3) * (1 / omega - 0.5)
