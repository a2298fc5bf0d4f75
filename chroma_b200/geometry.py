"""QDP++ cb2 site layout, restated for host-side bookkeeping (numpy).

idx = cb*Vh + ((t*Lz+z)*Ly+y)*(Lx/2) + x/2, cb = (x+y+z+t)&1 -- the layout the reference's own Dslash assumes
(other_libs/cpp_wilson_dslash/lib/shift_table_scalar.cc:155-214).
"""
import numpy as np


def site_coords(L):
    """int array [V,4]: (x,y,z,t) of every cb2 site index."""
    Lx, Ly, Lz, Lt = (int(v) for v in L)
    V = Lx * Ly * Lz * Lt
    Vh, Lxh = V // 2, Lx // 2
    idx = np.arange(V)
    cb, r = idx // Vh, idx % Vh
    xh = r % Lxh
    r = r // Lxh
    y = r % Ly
    r = r // Ly
    z = r % Lz
    t = r // Lz
    x = 2 * xh + ((cb + y + z + t) & 1)
    return np.stack([x, y, z, t], axis=1)


def site_index(L, c):
    """cb2 index of coordinates c = [...,4] (vectorised)."""
    c = np.asarray(c)
    Lx, Ly, Lz, Lt = (int(v) for v in L)
    V = Lx * Ly * Lz * Lt
    x, y, z, t = c[..., 0] % Lx, c[..., 1] % Ly, c[..., 2] % Lz, c[..., 3] % Lt
    cb = (x + y + z + t) & 1
    return ((t * Lz + z) * Ly + y) * (Lx // 2) + x // 2 + cb * (V // 2)
