#!/usr/bin/env python
"""Generates tests/golden/stepfuncs.npz by running the UNMODIFIED reference (oracle/_ref,
ultranest/stepfuncs.pyx compiled with Cython) on the seeded cases of tests/stepfuncs_cases.py.

    python tests/golden/make_golden_stepfuncs.py

Needs /root/reference (or a built oracle/_ref); the fixture it writes travels with the repo.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import stepfuncs_cases as cases  # noqa: E402


def main():
    oracle.reference()
    import ultranest.stepfuncs as sf
    out = {}
    # within_unit_cube
    u = cases.cube_case(1, 500, 7)
    out["cube_mask"] = sf.within_unit_cube(u)
    # evolve_update
    c = cases.evolve_update_case(2, 1000)
    sr, bi = sf.evolve_prepare(c["searching_left"], c["searching_right"])
    out["prep_search_right"], out["prep_bisecting"] = sr, bi
    success = np.zeros(1000, dtype=bool)
    sf.evolve_update(c["acceptable"], c["Lnew"], c["Lmin"], sr, bi, c["currentt"], c["current_left"],
                     c["current_right"], c["searching_left"], c["searching_right"], success)
    for k in ("currentt", "current_left", "current_right", "searching_left", "searching_right"):
        out["upd_" + k] = c[k]
    out["upd_success"] = success
    # step_back
    c = cases.step_back_case(3, 300, 12)
    sf.step_back(c["Lmin"], c["allL"], c["generation"], c["currentt"])
    out["back_allL"], out["back_generation"], out["back_currentt"] = c["allL"], c["generation"], c["currentt"]
    # evolve (global np.random stream, like the reference)
    st = cases.evolve_state(4, 800, 6)
    np.random.seed(5)
    (_, (success, unew, pnew, Lnew), nc) = sf.evolve(cases.identity, cases.gauss_loglike(0.5, 0.1),
                                                     -8.0, **st)
    for k, val in st.items():
        out["evo_" + k] = val
    out["evo_success"], out["evo_unew"], out["evo_pnew"], out["evo_Lnew"], out["evo_nc"] = success, unew, pnew, Lnew, nc
    # update_vectorised_slice_sampler over several passes
    c = cases.slice_sampler_case(6, 256, 5, 6)
    loglike = cases.gauss_loglike(0.5, c["sigma"])
    popsize = 256
    allu, allL, v = c["allu"].copy(), c["allL"].copy(), c["v"]
    allp = np.full_like(allu, np.nan)
    tleft, tright = c["tleft"].copy(), c["tright"].copy()
    tlw, trw = tleft.copy(), tright.copy()
    worker_running = np.arange(popsize, dtype=np.int64)
    status = np.zeros(popsize, dtype=np.int64)
    disc = []
    for it in range(len(c["draws"])):
        t = tlw + (trw - tlw) * c["draws"][it]
        pu = allu[worker_running, :] + t.reshape((-1, 1)) * v[worker_running, :]
        pp = pu.copy()
        pL = loglike(pp)
        tleft, tright, worker_running, status, allu, allL, allp, nd = sf.update_vectorised_slice_sampler(
            t, tleft, tright, pL, pu, pp, worker_running, status, c["Lmin"], c["shrink"], allu, allL, allp, popsize)
        disc.append(nd)
        tlw, trw = tleft[worker_running], tright[worker_running]
        out["slice_worker_%d" % it] = worker_running.copy()
        out["slice_status_%d" % it] = status.copy()
    out["slice_allu"], out["slice_allL"], out["slice_allp"] = allu, allL, allp
    out["slice_tleft"], out["slice_tright"], out["slice_discarded"] = tleft, tright, np.array(disc)
    path = os.path.join(HERE, "stepfuncs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
