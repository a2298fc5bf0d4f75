"""CPU check of the chunked traversal order of the Dslash kernels (engine_impl.cuh::set_chunks, dslash.cuh::mrhs_site /
launch_site): a Python restatement of the chunk choice and of the site decode, checked for the properties the kernels
rely on -- every site of the box is visited exactly once, chunks tile the box, three slices of a chunk fit the budget,
and no admissible divisor pair has a smaller surface.  (The CUDA decode itself is exercised by
tests/test_gpu_parity.py::test_multi_rhs_operator_parity with a shrunken L2 budget.)"""
import itertools

import numpy as np
import pytest


def set_chunks(Lxh, Ly, nz, nb, csize, budget, y_chunks=True):
    """-> (zc_sites, dz, dy, ncy); zc_sites = 0: natural order."""
    rowb = Lxh * 12 * csize * nb
    if nz <= 0 or budget <= 0 or 3 * rowb * Ly * nz <= budget:
        return 0, nz, Ly, 1
    best, bz, by, bvol = 1e30, 1, 1, 0
    for dz in range(1, nz + 1):
        if nz % dz:
            continue
        for dy in range(1, Ly + 1):
            if Ly % dy or 3 * rowb * dz * dy > budget:
                continue
            f = (2.0 / dz if dz < nz else 0.0) + (2.0 / dy if dy < Ly else 0.0)
            vol = dz * dy
            if f < best - 1e-12 or (f < best + 1e-12 and vol > bvol):
                best, bz, by, bvol = f, dz, dy, vol
    if not y_chunks:
        by, bz = Ly, 1
        for dz in range(1, nz + 1):
            if nz % dz == 0 and 3 * rowb * Ly * dz <= budget:
                bz = dz
    return bz * by * Lxh, bz, by, Ly // by


def decode(local, Lxh, Ly, Lz, box, zc_sites, dz, dy, ncy):
    t0, nt, z0, nz = box
    if zc_sites:
        per = zc_sites * nt
        zc, rem = divmod(local, per)
        tt, w = divmod(rem, zc_sites)
        cz, cy = divmod(zc, ncy)
        zz, r2 = divmod(w, dy * Lxh)
        yy, xh = divmod(r2, Lxh)
        return (((t0 + tt) * Lz + z0 + cz * dz + zz) * Ly + cy * dy + yy) * Lxh + xh
    row = Lxh * Ly
    q, w = divmod(local, row)
    tt, zz = divmod(q, nz)
    return ((t0 + tt) * Lz + z0 + zz) * row + w


CASES = [  # Lxh, Ly, Lz, Lt, box (t0, nt, z0, nz), nb, budget
    (4, 4, 6, 6, (0, 6, 0, 6), 12, 250 << 10),
    (4, 4, 6, 6, (0, 6, 0, 6), 7, 40 << 10),
    (4, 8, 8, 8, (1, 6, 1, 6), 12, 300 << 10),          # the interior of a Z x T split lattice
    (3, 6, 4, 4, (0, 4, 0, 4), 5, 60 << 10),
    (24, 48, 48, 4, (0, 4, 0, 48), 12, 40 << 20),         # 48^3 x 12 right-hand sides, fp64: the production case
    (32, 64, 64, 2, (0, 2, 0, 64), 1, 40 << 20),          # 64^3 single right-hand side
]


@pytest.mark.parametrize("y_chunks", [True, False])
@pytest.mark.parametrize("case", CASES)
def test_chunked_order_is_a_permutation_of_the_box(case, y_chunks):
    Lxh, Ly, Lz, Lt, box, nb, budget = case
    t0, nt, z0, nz = box
    zc_sites, dz, dy, ncy = set_chunks(Lxh, Ly, nz, nb, 16, budget, y_chunks)
    n = Lxh * Ly * nz * nt
    if zc_sites:
        assert nz % dz == 0 and Ly % dy == 0 and ncy * dy == Ly and zc_sites == dz * dy * Lxh
        if y_chunks:                                         # (the z-only rule keeps one plane even if that exceeds the budget)
            assert 3 * Lxh * 12 * 16 * nb * dz * dy <= budget
    idx = np.array([decode(l, Lxh, Ly, Lz, box, zc_sites, dz, dy, ncy) for l in range(n)])
    want = sorted(((t * Lz + z) * Ly + y) * Lxh + x for t in range(t0, t0 + nt) for z in range(z0, z0 + nz) for y in range(Ly) for x in range(Lxh))
    assert sorted(idx.tolist()) == want
    if zc_sites:                                            # a warp's 32 consecutive sites stay inside one chunk-slice when 32 | chunk
        first = idx[:zc_sites]
        ts = set((i // (Lxh * Ly * Lz)) for i in first.tolist())
        assert len(ts) == 1                                  # the first zc_sites sites are ONE time slice of ONE chunk


def test_production_chunk_shapes():
    # 48^3, 12 sources, fp64, 40 MB: 12 z-planes x 16 y-rows (surface 2/12 + 2/16 = 0.29 extra fetches per site; z-only: 4 planes, 0.5)
    zc, dz, dy, ncy = set_chunks(24, 48, 48, 12, 16, 40 << 20)
    assert (dz, dy, ncy, zc) == (12, 16, 3, 12 * 16 * 24)
    zc0, dz0, dy0, _ = set_chunks(24, 48, 48, 12, 16, 40 << 20, y_chunks=False)
    assert (dz0, dy0) == (4, 48)
    # 48^3 single source: three whole slices fit -> natural order
    assert set_chunks(24, 48, 48, 1, 16, 40 << 20)[0] == 0
    # 64^3 single source: 32 z-planes, y not split (same as the z-only rule)
    assert set_chunks(32, 64, 64, 1, 16, 40 << 20)[1:3] == (32, 64)
    # every admissible pair has a surface at least as large as the chosen one
    for (Lxh, Ly, nz, nb) in [(24, 48, 48, 12), (32, 64, 64, 12), (16, 32, 30, 12)]:
        zc, dz, dy, _ = set_chunks(Lxh, Ly, nz, nb, 16, 40 << 20)
        f = lambda a, b: (2.0 / a if a < nz else 0.0) + (2.0 / b if b < Ly else 0.0)
        for a, b in itertools.product(range(1, nz + 1), range(1, Ly + 1)):
            if nz % a == 0 and Ly % b == 0 and 3 * Lxh * 12 * 16 * nb * a * b <= (40 << 20):
                assert f(dz, dy) <= f(a, b) + 1e-12
