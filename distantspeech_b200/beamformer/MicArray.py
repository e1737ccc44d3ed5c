"""Array geometry, TDOAs and steering vectors -- host-side precompute mirroring
``DistantSpeech/beamformer/MicArray.py`` (MicArray :20, steering_vector :74,
compute_tau :96 / :149).  The reference builds a pyroomacoustics room on every
instantiation (:41); room simulation is data generation and out of scope here,
so ``array_sim`` is ``None``.
"""
import numpy as np


def cart2sph(x, y, z):
    azimuth = np.arctan2(y, x)
    elevation = np.arctan2(z, np.sqrt(x ** 2 + y ** 2))
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    return azimuth, elevation, r


def sph2cart(azimuth, elevation, r):
    x = r * np.cos(elevation) * np.cos(azimuth)
    y = r * np.cos(elevation) * np.sin(azimuth)
    z = r * np.sin(elevation)
    return x, y, z


def _tau_from_geometry(mic_loc, c, incident_angle, dist):
    """tau_m = -|r_m| cos(theta_m) / c with theta_m the angle between -r_m and the
    propagation direction (MicArray.py:120-143, :170-185).  ``dist`` is the
    source distance the reference uses to build the direction vector (10 in the
    method, 1 in the module function) -- it cancels except for rounding."""
    incident_angle = np.asarray(incident_angle, dtype=np.float64)
    az = incident_angle[0]
    el = incident_angle[1] if len(incident_angle.shape) > 0 else 0
    x0, y0, z0 = sph2cart(az, el, dist)
    p0 = -1 * np.array([x0, y0, z0])
    p0_norm = np.sqrt(np.sum(p0 * p0))
    M = mic_loc.shape[0]
    tau = np.zeros((M, 1))
    for m in range(M):
        v = -1 * mic_loc[m, :]
        v_norm = np.sqrt(np.sum(v * v))
        cos_theta = np.sum(v * p0) / (p0_norm * v_norm + 1e-12)
        tau[m] = -1 * v_norm * cos_theta / c
    return tau


class MicArray(object):
    def __init__(self, arrayType='circular', r=0.032, c=343, M=4, n_fft=256, energy_absorption=0.7,
                 room_size=[5.0, 3.0, 3.0]):
        self.arrayType = arrayType
        self.c = c
        self.r = r
        self.fs = 16000                      # hard-wired in the reference (MicArray.py:27)
        self.M = M
        self.n_fft = n_fft
        self.half_bin = round(self.n_fft / 2 + 1)
        self.freq_bin = np.linspace(0, self.half_bin - 1, self.half_bin)
        self.gamma = np.arange(0, 360, int(360 / self.M)) * np.pi / 180
        self.tau = np.zeros((self.M, 1))
        self.omega = 2 * np.pi * self.freq_bin * self.fs / self.n_fft
        self.array_type = arrayType
        self.mic_loc = np.zeros((M, 3))
        self.mic_loc = self.array_init()
        self.array_sim = None

    def array_init(self, mic_loc=None):
        if self.array_type == 'circular':
            az = np.arange(0, 360, int(360 / self.M)) * np.pi / 180
            for m in range(self.M):
                self.mic_loc[m, :] = sph2cart(az[m], 0, self.r)
        elif self.array_type == 'linear':
            self.mic_loc[:, 0] = -(np.arange(self.M) - (self.M - 1) / 2) * self.r
        else:
            assert self.mic_loc.shape == mic_loc.shape, \
                'user defined mic location should be 2-D array with shape M X 3'
            self.mic_loc = mic_loc
        return self.mic_loc

    def compute_tau(self, incident_angle, normalize=False):
        """Delay of each mic relative to the origin, [M, 1]; ``incident_angle`` in radians."""
        self.tau = _tau_from_geometry(self.mic_loc, self.c, incident_angle, 10)
        if normalize:
            self.tau = self.tau - self.tau[0, 0]
        return self.tau

    def steering_vector(self, look_direction=0):
        """Azimuth-only steering vectors [half_bin, M] (look_direction in degrees, :74-94)."""
        tau = self.compute_tau(incident_angle=np.array([look_direction, 0]) * np.pi / 180)
        return np.exp(-1j * self.omega[:, None] * tau[:, 0][None, :])


def compute_tau(mic_array: MicArray, incident_angle):
    """Module-level variant (MicArray.py:149-187), returns a fresh [M, 1] array."""
    return _tau_from_geometry(mic_array.mic_loc, mic_array.c, incident_angle, 1)
