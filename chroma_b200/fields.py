"""Synthetic lattice fields in QDP++ host layout (numpy), seeded and reproducible.

Recipes follow the reference's own tests: Gaussian 3x3 complex matrices re-unitarised by Gram-Schmidt
(other_libs/cpp_wilson_dslash/tests/testDslashFull.cc:43-47 + tests/reunit.cc; mainprogs/tests/symm_prec_tests.cc:41-45),
Gaussian fermion sources (testDslashFull.cc:50-52), and a smooth "weak field" stand-in for CFG_TYPE_WEAK_FIELD
(tests/chroma/hadron/propagator/prec_clover.ini.xml) for solver tests: Gram-Schmidt of 1 + eps*G.

Array conventions: gauge [4, V, 3, 3, 2], fermion [V, 4, 3, 2], site index = cb2 (see oracle/oracle.py).
Fields are generated directly in cb2 site order from one stream per (seed, mu), so the field at a given
site does not depend on how the lattice is later split across ranks.
"""
import numpy as np


def volume(L):
    return int(np.prod(L))


def _reunit(m):
    """Gram-Schmidt rows 0,1; row 2 = conj(row0 x row1).  m: complex [..., 3, 3]."""
    r0 = m[..., 0, :]
    r0 = r0 / np.linalg.norm(r0, axis=-1, keepdims=True)
    r1 = m[..., 1, :]
    r1 = r1 - r0 * np.sum(np.conj(r0) * r1, axis=-1, keepdims=True)
    r1 = r1 / np.linalg.norm(r1, axis=-1, keepdims=True)
    r2 = np.conj(np.cross(r0, r1))
    return np.stack([r0, r1, r2], axis=-2)


def _to_real(m):
    return np.ascontiguousarray(np.stack([m.real, m.imag], axis=-1))


def random_gauge(L, seed=11, dtype=np.float64):
    """Random SU(3) links (strong coupling: the hardest case for parity, a poorly conditioned one for solvers)."""
    V = volume(L)
    out = np.empty((4, V, 3, 3, 2), dtype=dtype)
    for mu in range(4):
        rng = np.random.default_rng([seed, mu])
        g = rng.standard_normal((V, 3, 3)) + 1j * rng.standard_normal((V, 3, 3))
        out[mu] = _to_real(_reunit(g))
    return out


def weak_gauge(L, seed=11, eps=0.2, dtype=np.float64):
    """Smooth field near the identity: reunitarise(1 + eps*G)."""
    V = volume(L)
    out = np.empty((4, V, 3, 3, 2), dtype=dtype)
    eye = np.eye(3, dtype=np.complex128)
    for mu in range(4):
        rng = np.random.default_rng([seed, mu, 7])
        g = rng.standard_normal((V, 3, 3)) + 1j * rng.standard_normal((V, 3, 3))
        out[mu] = _to_real(_reunit(eye + eps * g))
    return out


def unit_gauge(L, dtype=np.float64):
    V = volume(L)
    out = np.zeros((4, V, 3, 3, 2), dtype=dtype)
    for c in range(3):
        out[:, :, c, c, 0] = 1.0
    return out


def gaussian_fermion(L, seed=12, dtype=np.float64, cb=None):
    """Gaussian spinor on the full lattice; cb=0/1 zeroes the other checkerboard (gaussian(b, rb[cb]))."""
    V = volume(L)
    rng = np.random.default_rng([seed, 99])
    f = rng.standard_normal((V, 4, 3, 2)).astype(dtype)
    if cb is not None:
        Vh = V // 2
        f[(1 - cb) * Vh:(2 - cb) * Vh] = 0
    return f


def point_source(L, spin, colour, dtype=np.float64):
    """delta_{x,0} delta_{s,spin} delta_{c,colour} (site 0 = origin, even checkerboard, cb2 index 0)."""
    f = np.zeros((volume(L), 4, 3, 2), dtype=dtype)
    f[0, spin, colour, 0] = 1.0
    return f


def apply_bc(L, u, boundary=(1, 1, 1, -1)):
    """u[m] *= boundary[m] on the last slice of direction m (SimpleFermBC::modify, simple_fermbc.h:87-103)."""
    from .geometry import site_coords
    u = u.copy()
    c = site_coords(L)
    for mu in range(4):
        if boundary[mu] != 1:
            u[mu, c[:, mu] == L[mu] - 1] *= boundary[mu]
    return u


def random_clover(L, seed=13, dtype=np.float64, diag=4.0):
    """A random Hermitian, diagonally dominant PrimitiveClovTriang field [V,72] (as testClover.cc:118-150 does
    with drand48): used to test the clover apply / LDL^dagger inverse independently of the gauge field."""
    V = volume(L)
    rng = np.random.default_rng([seed, 5])
    tri = (0.3 * rng.standard_normal((V, 72))).astype(dtype)
    tri[:, :12] += diag
    return tri


def t_slab(full, L, t0, t1):
    """Rows of a full-lattice cb2-ordered array [V, ...] that belong to the time slab [t0, t1), in the LOCAL cb2 order
    of a rank that owns that slab (t is the slowest index inside each checkerboard, so each half is contiguous)."""
    V = volume(L)
    Vh = V // 2
    s3h = Vh // L[3]
    return np.concatenate([full[cb * Vh + t0 * s3h: cb * Vh + t1 * s3h] for cb in range(2)], axis=0)


def sub_lattice(full, L, lo, hi):
    """Rows of a full-lattice cb2-ordered array [V, ...] that belong to the box lo[mu] <= x_mu < hi[mu], in the LOCAL cb2
    order of a rank that owns that box (box extents even and lo even in every direction, so local parity == global
    parity).  Generalises t_slab to the T x Z process grids."""
    from .geometry import site_coords, site_index
    lo = np.asarray(lo)
    ext = tuple(int(h - l) for l, h in zip(lo, hi))
    assert all(e % 2 == 0 and e >= 2 for e in ext) and all(int(l) % 2 == 0 for l in lo)
    c = site_coords(ext) + lo[None, :]
    return full[site_index(L, c)]


def grid_box(L, proc_grid, proc_coord):
    """(lo, hi) of the sub-lattice of rank proc_coord in a proc_grid decomposition of L."""
    ext = [L[m] // proc_grid[m] for m in range(4)]
    lo = [proc_coord[m] * ext[m] for m in range(4)]
    return lo, [lo[m] + ext[m] for m in range(4)]
