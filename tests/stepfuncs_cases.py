"""Seeded input generators for the population step-sampler helpers (ultranest/stepfuncs.pyx),
shared by the golden-vector generator (tests/golden/make_golden_stepfuncs.py, runs the
reference), the CPU oracle tests and the GPU parity tests."""
import numpy as np


def gauss_loglike(centre=0.5, sigma=0.1):
    """docs/gauss.py:25-27 shape: the NumPy expression the device likelihood is bit-identical to."""
    def loglike(theta):
        d = theta.shape[1]
        return -0.5 * (((theta - centre) / sigma)**2).sum(axis=1) - 0.5 * np.log(2 * np.pi * sigma**2) * d
    return loglike


def identity(u):
    return u


def cube_case(seed, n, d):
    """Rows inside, on and outside the unit cube boundary, plus NaN."""
    rng = np.random.RandomState(seed)
    u = rng.uniform(0.001, 0.999, size=(n, d))
    bad = rng.uniform(size=n) < 0.4
    col = rng.randint(d, size=n)
    special = np.array([0.0, 1.0, -1e-300, 1.0 + 2.3e-16, np.nan, -0.5, 1.5, np.nextafter(1.0, 0),
                        np.nextafter(0.0, 1), np.inf])
    vals = special[rng.randint(len(special), size=n)]
    u[bad, col[bad]] = vals[bad]
    return u


def evolve_state(seed, n, d, scale=0.05):
    """Walker population in a mix of the three slice states (stepping out left / right,
    bisecting) with brackets that partly leave the unit cube."""
    rng = np.random.RandomState(seed)
    currentu = rng.uniform(0.3, 0.7, size=(n, d))
    v = rng.normal(size=(n, d))
    currentv = v / np.sqrt((v**2).sum(axis=1, keepdims=True)) * scale
    state = rng.randint(3, size=n)
    searching_left = state == 0
    searching_right = state <= 1
    searching_right[rng.uniform(size=n) < 0.1] ^= True
    grow = 2.0**rng.randint(0, 6, size=n)
    current_left = -1.0 * grow
    current_right = 1.0 * 2.0**rng.randint(0, 6, size=n)
    currentt = np.where(rng.uniform(size=n) < 0.5, -1.0, 1.0) * rng.uniform(size=n)
    currentt[rng.uniform(size=n) < 0.05] = 0.0
    currentL = rng.normal(size=n)
    return dict(currentu=currentu, currentL=currentL, currentt=currentt, currentv=currentv,
                current_left=current_left, current_right=current_right,
                searching_left=searching_left, searching_right=searching_right)


def evolve_update_case(seed, n):
    rng = np.random.RandomState(seed)
    st = evolve_state(seed + 1, n, 2)
    acceptable = rng.uniform(size=n) < 0.7
    Lmin = 0.25
    Lnew = rng.normal(size=int(acceptable.sum())) + 0.25
    if len(Lnew) > 4:
        Lnew[0] = Lmin               # `>` edge
        Lnew[1] = np.nan
        Lnew[2] = np.nextafter(Lmin, 1)
        Lnew[3] = -np.inf
    return dict(acceptable=acceptable, Lnew=Lnew, Lmin=Lmin, currentt=st["currentt"],
                current_left=st["current_left"], current_right=st["current_right"],
                searching_left=st["searching_left"], searching_right=st["searching_right"])


def step_back_case(seed, n, nsteps):
    """Chains of different lengths (NaN beyond the walker's generation) whose earlier entries
    partly fall below the raised threshold."""
    rng = np.random.RandomState(seed)
    generation = rng.randint(-1, nsteps + 1, size=n).astype(np.int64)
    allL = np.full((n, nsteps + 1), np.nan)
    for i in range(n):
        g = generation[i]
        if g >= 0:
            allL[i, : g + 1] = np.sort(rng.normal(size=g + 1)) if rng.uniform() < 0.5 else rng.normal(size=g + 1)
    Lmin = -0.3
    allL[rng.randint(n), 0] = Lmin   # equal is not below
    currentt = rng.normal(size=n)
    return dict(Lmin=Lmin, allL=allL, generation=generation, currentt=currentt)


def slice_sampler_case(seed, popsize, d, iters, shrink=1.0, sigma=0.05):
    """Start state of PopulationSimpleSliceSampler's inner loop (popstepsampler.py:916-938) and
    the uniform draws of `iters` passes."""
    rng = np.random.RandomState(seed)
    allu = rng.uniform(0.35, 0.65, size=(popsize, d))
    v = rng.normal(size=(popsize, d))
    v = v / np.sqrt((v**2).sum(axis=1, keepdims=True)) * 0.3
    # unit-cube intersections of the slices (popstepsampler.py:26-61)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (0.0 - allu) / v
        t2 = (1.0 - allu) / v
    tleft = np.max(np.where(np.minimum(t1, t2) < 0, np.minimum(t1, t2), -np.inf), axis=1)
    tright = np.min(np.where(np.maximum(t1, t2) > 0, np.maximum(t1, t2), np.inf), axis=1)
    loglike = gauss_loglike(0.5, sigma)
    allL = loglike(allu)
    Lmin = float(np.median(allL))
    draws = rng.uniform(size=(iters, popsize))
    return dict(allu=allu, allL=allL, v=v, tleft=tleft, tright=tright, Lmin=Lmin,
                shrink=shrink, draws=draws, sigma=sigma)
