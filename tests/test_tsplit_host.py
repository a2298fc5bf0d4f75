"""CPU (gloo, world_size 2) tests of the host-side logic of the multi-GPU path: the b200_comm bootstrap callbacks that
carry the CUDA IPC handles, and the T-slab decomposition of cb2-ordered fields.  No GPU, no engine call."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import make_comm
    comm = make_comm(dist, rank, world)
    # what Halo::init does with the comm: allgather one 64-byte "IPC handle" per rank, then barrier
    send = (C.c_ubyte * 64)(*[(rank * 37 + i) % 251 for i in range(64)])
    recv = (C.c_ubyte * (64 * world))()
    rc = comm.allgather(None, C.cast(send, C.c_void_p), C.cast(recv, C.c_void_p), 64)
    ok = rc == 0 and comm.barrier(None) == 0
    got = np.frombuffer(recv, dtype=np.uint8).reshape(world, 64)
    for r in range(world):
        ok &= bool((got[r] == np.array([(r * 37 + i) % 251 for i in range(64)], dtype=np.uint8)).all())
    ok &= comm.rank == rank and comm.size == world
    q.put((rank, ok))
    dist.destroy_process_group()


def test_comm_bootstrap_callbacks_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


@pytest.mark.parametrize("latt,world", [((4, 4, 4, 8), 2), ((6, 4, 2, 8), 4), ((4, 4, 4, 4), 2)])
def test_t_slab_is_local_cb2_order(latt, world):
    """A slab cut out of a global cb2-ordered field is exactly the field a rank would hold in its LOCAL cb2 order
    (local extents even => local parity == global parity), and the slabs tile the lattice."""
    from chroma_b200 import fields, geometry
    V = int(np.prod(latt))
    c = geometry.site_coords(latt)                       # global coordinates per global index
    lt = latt[3] // world
    seen = np.zeros(V, dtype=int)
    for r in range(world):
        t0, t1 = r * lt, (r + 1) * lt
        loc = fields.t_slab(c, latt, t0, t1)             # coordinates carried along by the slicing
        llatt = (latt[0], latt[1], latt[2], lt)
        want = geometry.site_coords(llatt)               # local coordinates in local cb2 order
        want[:, 3] += t0
        assert np.array_equal(loc, want)
        seen[geometry.site_index(latt, loc)] += 1
    assert (seen == 1).all()


def test_halo_protocol_on_cpu(oracle):
    """Restate the face-exchange protocol of halo.cuh with numpy on T slabs and check that it reproduces the global
    hopping term on every rank:  a rank computes all hops from LOCAL data, except that on its last time slice the
    forward hop uses ghost_fwd = (1 - s g3) psi sent by the +t rank (projected only; the receiver multiplies by its own
    U_t), and on its first slice the backward hop uses ghost_bwd = U_t^dag (1 + s g3) psi computed by the -t rank from
    ITS links (sender multiplies: decomp_hvv, cpp_dslash_parscalar_utils_64bit.cc:61-110)."""
    from chroma_b200 import fields, geometry
    from test_oracle import GAMMA, cplx
    latt, world = (4, 4, 2, 8), 4
    g = oracle.Geom(latt)
    u = fields.apply_bc(latt, fields.random_gauge(latt, seed=3))
    psi = fields.gaussian_fermion(latt, seed=4)
    pk = oracle.pack_gauge(latt, u)
    c = geometry.site_coords(latt)
    U, P = cplx(u), cplx(psi)
    lt = latt[3] // world
    for s in (+1, -1):
        want = cplx(oracle.dslash(g, psi, pk, s, 0) + oracle.dslash(g, psi, pk, s, 1))
        for r in range(world):
            t0, t1 = r * lt, (r + 1) * lt
            mine = np.where((c[:, 3] >= t0) & (c[:, 3] < t1))[0]
            owned = np.zeros(len(c), dtype=bool)
            owned[mine] = True
            got = np.zeros((len(mine), 4, 3), dtype=complex)
            for mu in range(4):
                e = np.zeros(4, dtype=int)
                e[mu] = 1
                Pm, Pp = np.eye(4) - s * GAMMA[mu], np.eye(4) + s * GAMMA[mu]
                f = geometry.site_index(latt, c[mine] + e)
                b = geometry.site_index(latt, c[mine] - e)
                # half spinors as whoever OWNS the neighbour site computes them
                hf = np.einsum("st,xtc->xsc", Pm, P[f])                          # projected by the owner of f
                hb = np.einsum("xba,xsb->xsa", U[mu][b].conj(), np.einsum("st,xtc->xsc", Pp, P[b]))   # projected AND multiplied by the owner of b
                if mu < 3:
                    assert owned[f].all() and owned[b].all()                     # spatial hops never leave the slab
                else:
                    # exactly the boundary slices need a neighbour rank; the link U_t(b) is NOT local there
                    assert (~owned[f]).sum() == (c[mine, 3] == t1 - 1).sum() and (~owned[b]).sum() == (c[mine, 3] == t0).sum()
                got += np.einsum("xab,xsb->xsa", U[mu][mine], hf) + hb           # receiver multiplies the forward hop by ITS link
            assert np.abs(got - want[mine]).max() < 1e-13


@pytest.mark.parametrize("latt,grid", [((4, 4, 8, 8), (1, 1, 2, 2)), ((4, 6, 4, 8), (1, 1, 2, 4)), ((4, 4, 8, 4), (1, 1, 4, 1))])
def test_sub_lattice_is_local_cb2_order(latt, grid):
    """T x Z boxes cut out of a global cb2-ordered field are in the LOCAL cb2 order of their rank, and tile the lattice."""
    from chroma_b200 import fields, geometry
    V = int(np.prod(latt))
    c = geometry.site_coords(latt)
    seen = np.zeros(V, dtype=int)
    for pt in range(grid[3]):
        for pz in range(grid[2]):
            lo, hi = fields.grid_box(latt, grid, (0, 0, pz, pt))
            loc = fields.sub_lattice(c, latt, lo, hi)
            want = geometry.site_coords(tuple(h - l for l, h in zip(lo, hi))) + np.asarray(lo)[None, :]
            assert np.array_equal(loc, want)
            seen[geometry.site_index(latt, loc)] += 1
    assert (seen == 1).all()


def test_z_face_index_matches_on_both_sides():
    """halo.cuh addresses a Z face by f = (t*Ly + y)*Lxh + xh on BOTH sides of the cut: the sender packs its z = 0 (or
    Lz-1) site of the source parity into slot f, and the receiver's target site at z = Lz-1 (or 0) reads slot f computed
    from its own (xh, y, t).  Check with the cb2 index arithmetic that the two are the same physical neighbour pair."""
    from chroma_b200 import geometry
    ll = (6, 4, 4, 2)                                     # local extents of every rank (even)
    Lxh, Ly, Lz, Lt = ll[0] // 2, ll[1], ll[2], ll[3]
    Vh = int(np.prod(ll)) // 2
    c = geometry.site_coords(ll)
    row = Lxh * Ly
    for par in (0, 1):                                    # target parity
        tgt = c[par * Vh:(par + 1) * Vh]
        idx = np.arange(Vh)
        xh = idx % Lxh
        for zface, znbr in ((Lz - 1, 0), (0, Lz - 1)):    # forward hop from z = Lz-1 reads the +z rank's z = 0 plane, and v.v.
            m = tgt[:, 2] == zface
            f_recv = (tgt[m, 3] * Ly + tgt[m, 1]) * Lxh + xh[m]
            # the neighbour site as the OTHER rank indexes it: same (x, y, t), z = znbr, parity 1 - par
            nb = tgt[m].copy()
            nb[:, 2] = znbr
            assert ((nb.sum(axis=1) & 1) == 1 - par).all()
            nidx = geometry.site_index(ll, nb) - (1 - par) * Vh
            t_s, w_s = nidx // (row * Lz), nidx % row
            assert np.array_equal(nidx, t_s * row * Lz + znbr * row + w_s)      # the sender's idx formula in pack_faces_kernel
            assert np.array_equal(t_s * row + w_s, f_recv)


def test_site_boxes_tile_the_local_lattice():
    """Engine::launch_dslash splits a T x Z-split local lattice into an interior box and up to four boundary boxes;
    restate box_site (common.cuh) and check that the boxes tile one checkerboard exactly once, in every split mode."""
    def box_site(g, b, local):
        Lxh, Ly, Lz, Lt = g
        t0, nt, z0, nz = b
        row = Lxh * Ly
        if nz == Lz:
            return t0 * row * Lz + local
        w, q = local % row, local // row
        return ((t0 + q // nz) * Lz + z0 + q % nz) * row + w

    for g in ((2, 4, 6, 8), (2, 2, 2, 4), (3, 2, 4, 2), (2, 2, 2, 2)):
        Lxh, Ly, Lz, Lt = g
        Vh = Lxh * Ly * Lz * Lt
        for tsplit in (0, 1):
            for zsplit in (0, 1):
                if not (tsplit or zsplit):
                    continue
                inner = (1 if tsplit else 0, Lt - 2 if tsplit else Lt, 1 if zsplit else 0, Lz - 2 if zsplit else Lz)
                boxes = [inner] if inner[1] * inner[3] > 0 else []
                if tsplit:
                    boxes += [(0, 1, 0, Lz), (Lt - 1, 1, 0, Lz)]
                if zsplit and inner[1] > 0:
                    boxes += [(inner[0], inner[1], 0, 1), (inner[0], inner[1], Lz - 1, 1)]
                seen = np.zeros(Vh, dtype=int)
                for b in boxes:
                    for local in range(Lxh * Ly * b[3] * b[1]):
                        seen[box_site(g, b, local)] += 1
                assert (seen == 1).all(), (g, tsplit, zsplit)
