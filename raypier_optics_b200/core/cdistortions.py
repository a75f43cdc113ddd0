"""Host mirror of raypier/core/cdistortions.pyx (parameter holders; the Zernike
sums run on the GPU from a host-built evaluation tape, see scene.build_zernike_tapes)."""
import math

from .ctracer import Distortion

jnm_map = [(0, 0, 0)]


def eval_nmk(j):
    """ANSI single index j -> (n, m, k) (cdistortions.pyx:104-136)."""
    if j < 0:
        raise ValueError("J must be non-negative")
    if j >= len(jnm_map):
        n, m, k = jnm_map[-1]
        while True:
            while m < n:
                m += 2
                next_j = (n * (n + 2) + m) // 2
                half_n = int(math.floor(n / 2.))
                k = half_n * (half_n + 1) + abs(m)
                jnm_map.append((n, m, k))
                if next_j == j:
                    return (n, m, k)
                elif next_j > j:
                    raise ValueError("Something went wrong here! Missed j value.")
            n += 1
            m = -2 - n
    return jnm_map[j]


class SimpleTestZernikeJ7(Distortion):
    """cdistortions.pyx:39-81"""

    def __init__(self, **kwds):
        self.unit_radius = kwds.get("unit_radius", 1.0)
        self.amplitude = kwds.get("amplitude", 1.0)


class ZernikeDistortion(Distortion):
    """cdistortions.pyx:323-514.  ``ZernikeDistortion(unit_radius=10., j4=2e-3, j7=1e-3)``
    or ``ZernikeDistortion([(4, 2e-3), (7, 1e-3)], unit_radius=10.)``."""

    def __init__(self, *args, **coefs):
        self.coef_map = {}
        self.unit_radius = coefs.get("unit_radius", 1.0)
        cdict = {int(k[1:]): float(v) for k, v in coefs.items() if k.startswith("j")}
        if args:
            for k, v in args[0]:
                cdict[int(k)] = float(v)
        self.k_max = 0
        self.set_coefs(list(cdict.items()))

    def set_coefs(self, coefs):
        clist = sorted(coefs)
        self.j_max = max(j for j, v in clist)
        self.n_coefs = len(clist)
        self._coefs = []
        n_max = 0
        self.coef_map = {}
        for i, (j, v) in enumerate(clist):
            n, m, k = eval_nmk(j)
            self.coef_map[j] = i
            self._coefs.append([j, n, m, k, v])
            if n > n_max:
                n_max = n
        k_max = (n_max // 2) * (n_max // 2 + 1) + n_max + 1
        if k_max > self.k_max:
            self.k_max = k_max

    def __getitem__(self, idx):
        if idx >= self.n_coefs:
            raise IndexError("Index %d greater than number of coeffs (%d)." % (idx, self.n_coefs))
        return tuple(self._coefs[idx])

    def __len__(self):
        return self.n_coefs

    def update_coef(self, j, value):
        self._coefs[self.coef_map[j]][4] = float(value)

    def get_coef(self, j):
        return self._coefs[self.coef_map[j]][4]
