"""Measure the SURVEY section-8 (f4) additions on one B200: the symmetric even-odd operator and the multi-shift CG,
48^3x96 fp64 (same synthetic field as bench.py).  Prints one JSON line; not the headline bench."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from chroma_b200 import lib as L
    from chroma_b200.solver import Context
    latt = tuple(int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (48, 48, 48, 96)
    n_shift = int(os.environ.get("N_SHIFT", "10"))
    dev = torch.device("cuda", 0)
    ctx = Context(latt, prec="double", device=0)
    u = bench.torch_weak_gauge(latt, 0, latt, 11, 0.2, dev)
    bench.apply_bc_local(u, latt, True)
    ctx.load_gauge(u, t_boundary=-1)
    del u
    ctx.make_clover(4.1, 0.5, 0.5)
    Vh = ctx.Vh
    chi = bench.torch_gaussian_source(latt, 0, 12, dev, torch.float64).numpy()
    peak, how = bench.peaks()
    out = {"lattice": list(latt), "peak_gbs": peak, "peak_source": how}
    chi_f, o_f, psi_f = ctx.field(chi), ctx.field(), ctx.field()

    def m_times(tag):
        res = {}
        for isign in (+1, -1):
            ctx.dev_time_matpc(o_f, chi_f, isign, 3)
            a, b = ctx.dev_time_matpc(o_f, chi_f, isign, 20)
            res["M" if isign > 0 else "Mdag"] = {"ms_kernel1": a, "ms_kernel2": b, "ms": a + b,
                                                 "gflops": bench.FLOP_M * Vh / ((a + b) * 1e-3) * 1e-9}
        out[tag] = res

    m_times("asymmetric")
    info = ctx.dev_invert(psi_f.zero(), chi_f, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=2000)
    out["asymmetric"]["cg_solve"] = {"iterations": info.n_count, "secs": info.secs, "rel_resid": info.rel_resid}
    info = ctx.dev_invert(psi_f.zero(), chi_f, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-8, max_iter=2000)
    out["asymmetric"]["bicgstab_solve"] = {"iterations": info.n_count, "secs": info.secs, "rel_resid": info.rel_resid}
    # mixed-precision reliable updates (fp32 recurrences, fp64 residual replacement), same target
    for name, fn in (("reliable_cg_solve", ctx.dev_invert_reliable), ("reliable_bicgstab_solve", ctx.dev_invert_reliable_bicgstab)):
        fn(psi_f.zero(), chi_f, rsd=1e-8, delta=0.1, max_iter=2000)          # first call builds the fp32 twin
        info = fn(psi_f.zero(), chi_f, rsd=1e-8, delta=0.1, max_iter=2000)
        out["asymmetric"][name] = {"iterations": info.n_count, "secs": info.secs, "rel_resid": info.rel_resid, "n_updates": info.n_updates}
    # multi-shift on the asymmetric operator
    shifts = list(np.geomspace(1e-4, 2.0, n_shift))
    mp = ctx.mfield(n_shift)
    infos = ctx.dev_invert_multishift(mp, chi_f, shifts, 1e-8, max_iter=4000)
    infos = ctx.dev_invert_multishift(mp, chi_f, shifts, 1e-8, max_iter=4000)
    R = 8
    bytes_iter = 2 * 4416 + (4 * n_shift + 3 + 3) * 24 * R     # 2 M + fused residual update + shift update pass, per odd site
    out["multishift"] = {"n_shift": n_shift, "shifts": [shifts[0], shifts[-1]], "iterations": infos[0].n_count, "secs": infos[0].secs,
                         "ms_per_iteration": infos[0].secs / max(1, infos[0].n_count) * 1e3,
                         "worst_rel_resid": max(i.rel_resid for i in infos), "converged": all(i.converged for i in infos),
                         "algorithmic_gbs": bytes_iter * Vh * infos[0].n_count / infos[0].secs * 1e-9,
                         "frac_of_peak": bytes_iter * Vh * infos[0].n_count / infos[0].secs * 1e-9 / peak}
    # the same shifts one at a time would need n_shift CG solves; time the hardest one (smallest shift ~ plain CG)
    del mp
    ctx.set_preconditioning(True)
    m_times("symmetric")
    info = ctx.dev_invert(psi_f.zero(), chi_f, solver=L.B200_SOLVER_CG, rsd=1e-8, max_iter=2000)
    out["symmetric"]["cg_solve"] = {"iterations": info.n_count, "secs": info.secs, "rel_resid": info.rel_resid}
    info = ctx.dev_invert(psi_f.zero(), chi_f, solver=L.B200_SOLVER_BICGSTAB, rsd=1e-8, max_iter=2000)
    out["symmetric"]["bicgstab_solve"] = {"iterations": info.n_count, "secs": info.secs, "rel_resid": info.rel_resid}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
