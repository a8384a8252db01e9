"""LISA noise curves and the signal-to-noise ratio (pttools/omgw0/noise.py:7-329; constants pttools/omgw0/const.py).

Closed forms on the <= 1000 frequencies of a spectrum: host numpy.  For sweeps the frequency grid f = y / r_star * f_*0
is shared by all points, so the noise curve is evaluated once here and the per-point reduction
sqrt(T_obs * trapezoid(signal^2 / noise^2, f)) runs on the device over the omgw0 table (ptt_snr_batch, snr_weights)."""
from __future__ import annotations

import numpy as np

C_LIGHT = 299792458.0                                  # const.py:4
F1_LISA = 4e-4                                         # const.py:16
PC_TO_M = 3.0857e16
H0_HZ = 67.4 * 1e3 / (PC_TO_M * 1e6)                   # const.py:26-27
LISA_ARM_LENGTH = 2.5e9                                # const.py:30
YEAR_IN_SECONDS = 365.2425 * 24 * 60 * 60
LISA_OBS_TIME = 4 * 0.75 * YEAR_IN_SECONDS             # const.py:36


def ft(L=LISA_ARM_LENGTH):
    """Transfer frequency c / (2 pi L) (noise.py:55-60)."""
    return C_LIGHT / (2 * np.pi * L)


FT_LISA = ft()
F2_LISA = 4 / 3 * FT_LISA


def N_acc(L=LISA_ARM_LENGTH):
    return (3e-15 / L) ** 2                            # noise.py:67-78


def S_I(f, L=LISA_ARM_LENGTH):
    return 4 * N_acc(L) * (1 + (F1_LISA / f) ** 2)     # noise.py:258-263


def P_acc(f, L=LISA_ARM_LENGTH):
    return S_I(f, L) / (4 * (2 * np.pi * f) ** 4)      # noise.py:148-155


def P_oms(L=LISA_ARM_LENGTH):
    return (1.5e-11 / L) ** 2                          # noise.py:158-170


def W(f, f_t):
    return 1 - np.exp(-2j * f / f_t)                   # noise.py:287-292


def N_AE(f, f_t=FT_LISA, L=LISA_ARM_LENGTH, W_abs2=None):
    """Noise of the A and E channels (noise.py:81-98)."""
    c = np.cos(f / f_t)
    if W_abs2 is None:
        W_abs2 = np.abs(W(f, f_t)) ** 2
    return ((4 + 2 * c) * P_oms(L) + 8 * (1 + c + c ** 2) * P_acc(f, L)) * W_abs2


def R_AE(f, f_t=FT_LISA, W_abs2=None):
    """Response of the A and E channels (noise.py:173-182)."""
    if W_abs2 is None:
        W_abs2 = np.abs(W(f, f_t)) ** 2
    return 9 / 20 * W_abs2 / (1 + (3 * f / (4 * f_t)) ** 2)


def R_LISA(f, f2=F2_LISA):
    return 1 + (f / f2) ** 2


def S(N, R):
    return N / R


def S_AE(f, f_t=FT_LISA, L=LISA_ARM_LENGTH, both_channels=True):
    """Noise power spectral density of the A and E channels (noise.py:202-218); |W|^2 cancels."""
    ret = S(N_AE(f, f_t, L, W_abs2=1), R_AE(f, f_t, W_abs2=1))
    return ret / np.sqrt(2) if both_channels else ret


def S_AE_approx(f, L=LISA_ARM_LENGTH, both_channels=True):
    ret = 40 / 3 * (P_oms(L) + 4 * P_acc(f, L)) * (1 + (3 * f / (4 * ft(L))) ** 2)        # noise.py:221-255
    return ret / np.sqrt(2) if both_channels else ret


def S_gb(f, A=9e-35, f_ref_gb=1, fk=1.13e-3, a=0.138, b=-221, c=521, d=1680):
    """Galactic binaries (noise.py:266-284)."""
    return A * (1e-3 / f) ** (-7 / 3) * np.exp(-(f / f_ref_gb) ** a - b * f * np.sin(c * f)) * (1 + np.tanh(d * (fk - f)))


def omega(f, S_):
    """Sensitivity S -> Omega = 4 pi^2 / (3 H0^2) f^3 S (noise.py:101-113)."""
    return 4 * np.pi ** 2 / (3 * H0_HZ ** 2) * f ** 3 * S_


def omega_eb(f, f_ref_eb=25, omega_ref_eb=8.9e-10):
    return omega_ref_eb * (f / f_ref_eb) ** (2 / 3)    # noise.py:116-122


def omega_gb(f):
    return omega(f, S_gb(f))


def omega_ins(f):
    return omega(f, S_AE(f))


def omega_noise(f):
    """Instrument + extragalactic + galactic binaries (noise.py:140-145)."""
    return omega_ins(f) + omega_eb(f) + omega_gb(f)


def signal_to_noise_ratio(f, signal, noise, f_noise=None, obs_time=LISA_OBS_TIME, f_min=None, f_max=None):
    """noise.py:7-52, including its index conventions (i_f_max = -1 drops the last sample when only f_min is given)."""
    f, signal, noise = np.asarray(f, dtype=float), np.asarray(signal, dtype=float), np.asarray(noise, dtype=float)
    if f_noise is None:
        if not (f_min is None and f_max is None):
            i0 = 0 if f_min is None else int(np.argmax(f >= f_min))
            i1 = -1 if f_max is None else int(np.argmax(f >= f_max))
            f, noise, signal = f[i0:i1], noise[i0:i1], signal[i0:i1]
    else:
        f_noise = np.asarray(f_noise, dtype=float)
        if f_min is None:
            f_min = max(f[0], f_noise[0])
        if f_max is None:
            f_max = min(f[-1], f_noise[-1])
        i0, i1 = int(np.argmax(f_noise >= f_min)), int(np.argmax(f_noise >= f_max))
        f_gw = f
        f = f_noise[i0:i1]
        noise = noise[i0:i1]
        signal = 10. ** np.interp(np.log10(f), np.log10(f_gw), np.log10(signal))
    return np.sqrt(obs_time * np.trapezoid(signal ** 2 / noise ** 2, f))


def snr_weights(f, noises, obs_time=LISA_OBS_TIME) -> np.ndarray:
    """Rows T_obs * trapezoid weight(f) / noise^2 for ptt_snr_batch: snr = sqrt(sum_k weight_k signal_k^2), the same sum
    as signal_to_noise_ratio(f, signal, noise) on the full grid."""
    f = np.asarray(f, dtype=float)
    w = np.empty_like(f)
    w[1:-1] = 0.5 * (f[2:] - f[:-2])
    w[0], w[-1] = 0.5 * (f[1] - f[0]), 0.5 * (f[-1] - f[-2])
    return np.stack([obs_time * w / np.asarray(n, dtype=float) ** 2 for n in noises])
