"""Implementation-agnostic checks of the population step-sampler helpers against the golden
vectors of the reference (tests/golden/stepfuncs.npz).  `impl` is any module-like object with the
reference's function names (oracle.stepport on the CPU tier, ultranest_b200.stepfuncs on the GPU
tier, the compiled reference itself when oracle/_ref is present)."""
import os

import numpy as np

import stepfuncs_cases as cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stepfuncs.npz")


def golden():
    return np.load(GOLDEN)


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.dtype.kind == "f":
        np.testing.assert_array_equal(a, b)   # NaN == NaN positionally, bit-equal otherwise
        assert (np.signbit(a) == np.signbit(b)).all()
    else:
        assert (a.astype(np.int64) == b.astype(np.int64)).all()


def check_within_unit_cube(impl, g):
    _same(impl.within_unit_cube(cases.cube_case(1, 500, 7)), g["cube_mask"])
    assert impl.within_unit_cube(np.empty((0, 3))).shape == (0,)
    assert impl.within_unit_cube(np.full((1, 1), 0.5)).all()


def check_evolve_update(impl, g):
    c = cases.evolve_update_case(2, 1000)
    sr, bi = impl.evolve_prepare(c["searching_left"], c["searching_right"])
    _same(sr, g["prep_search_right"])
    _same(bi, g["prep_bisecting"])
    success = np.zeros(1000, dtype=bool)
    impl.evolve_update(c["acceptable"], c["Lnew"], c["Lmin"], sr, bi, c["currentt"], c["current_left"],
                       c["current_right"], c["searching_left"], c["searching_right"], success)
    for k in ("currentt", "current_left", "current_right", "searching_left", "searching_right"):
        _same(c[k], g["upd_" + k])
    _same(success, g["upd_success"])


def check_step_back(impl, g):
    c = cases.step_back_case(3, 300, 12)
    impl.step_back(c["Lmin"], c["allL"], c["generation"], c["currentt"])
    _same(c["allL"], g["back_allL"])
    _same(c["generation"], g["back_generation"])
    _same(c["currentt"], g["back_currentt"])
    # nothing below the threshold: untouched
    c = cases.step_back_case(3, 50, 4)
    before = c["allL"].copy()
    impl.step_back(-1e300, c["allL"], c["generation"], c["currentt"])
    _same(c["allL"], before)


def check_evolve(impl, g, transform=cases.identity, loglike=None):
    st = cases.evolve_state(4, 800, 6)
    np.random.seed(5)
    loglike = loglike or cases.gauss_loglike(0.5, 0.1)
    (ret_state, (success, unew, pnew, Lnew), nc) = impl.evolve(transform, loglike, -8.0, **st)
    for k, val in st.items():
        _same(val, g["evo_" + k])
    _same(success, g["evo_success"])
    _same(unew, g["evo_unew"])
    _same(pnew, g["evo_pnew"])
    _same(Lnew, g["evo_Lnew"])
    assert nc == int(g["evo_nc"])
    # the returned state objects are the caller's arrays (in-place contract, stepfuncs.pyx:246-248)
    assert ret_state[0] is st["currentt"] and ret_state[2] is st["current_left"]
    assert 0 < success.sum() < len(success)


def check_slice_sampler(impl, g):
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
        tleft, tright, worker_running, status, allu, allL, allp, nd = impl.update_vectorised_slice_sampler(
            t, tleft, tright, pL, pu, pp, worker_running, status, c["Lmin"], c["shrink"], allu, allL, allp, popsize)
        disc.append(nd)
        tlw, trw = tleft[worker_running], tright[worker_running]
        _same(worker_running, g["slice_worker_%d" % it])
        _same(status, g["slice_status_%d" % it])
    _same(allu, g["slice_allu"])
    _same(allL, g["slice_allL"])
    _same(allp, g["slice_allp"])
    _same(tleft, g["slice_tleft"])
    _same(tright, g["slice_tright"])
    _same(np.array(disc), g["slice_discarded"])
    assert 0 < (status == 0).sum() < popsize or (status == 1).all()


ALL_CHECKS = [check_within_unit_cube, check_evolve_update, check_step_back, check_evolve,
              check_slice_sampler]
