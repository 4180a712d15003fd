"""MAP_STATS (MAP_MODULE.F:182-303) and CALC_PERCENTILE (OUTPUT_MODULE.F:53-124) on the device (csrc/stats.cu) against a
plain numpy restatement of the two Fortran routines: through the CPU emulation of the kernel sources here, through
libapmlcl.so on the GPU with -m gpu."""
import numpy as np
import pytest

from helpers import load_map, synth_aperture
from apmapflow_b200 import lcl


def ref_percentiles(values, pct):
    """Sorted, duplicates collapsed, first distinct value whose cumulative count exceeds ceil(p/100 n) (the last one if
    none does): OUTPUT_MODULE.F:88-121."""
    s = np.sort(np.asarray(values, dtype=np.float64).ravel())
    vals, counts = np.unique(s, return_counts=True)
    cum = np.cumsum(counts)
    out = []
    for p in pct:
        k = np.ceil(p / 100.0 * s.size)
        j = np.searchsorted(cum, k, side="right")          # first index with cum > k
        out.append(vals[min(j, len(vals) - 1)])
    return np.array(out)


def ref_map_stats(m):
    """MAP_MODULE.F:218-291, loops as written (numpy sums; the correlation lengths are integer lags)."""
    nz, nx = m.shape
    n = nx * nz
    avg = m.sum() / n
    sm = (m - avg).sum() ** 2 / n
    var = (((m - avg) ** 2).sum() - sm) / (n - 1)
    res = {"avg": avg, "rms": np.sqrt((m ** 2).sum() / n), "dev": np.sqrt(var)}
    kmax = int(min(nx, nz) / 2.0) + 1
    for name, axis in (("x_corr", 1), ("z_corr", 0)):
        t1 = t2 = t3 = c = 0.0
        for k in range(1, kmax + 1):
            d = (m[:, :-k] - m[:, k:]) if axis == 1 else (m[:-k, :] - m[k:, :])
            if d.size == 0:
                continue
            with np.errstate(invalid="ignore", divide="ignore"):
                v3 = (d ** 2).sum() / d.size
                c = float(k)
                t1 = abs(v3 / var - 0.90)
                if v3 / var > 0.90:
                    if t2 < t1:
                        c = t3
                    break
            t2, t3 = t1, c
        res[name] = c
    return res


PCT = [0, 1, 5, 10, 15, 20, 25, 30, 40, 50, 60, 70, 75, 80, 85, 90, 95, 99, 100]      # the reference's input templates


def check_percentiles(lib, values, pct=PCT):
    got = lcl.percentiles(values, pct, lib=lib)
    assert np.array_equal(got, ref_percentiles(values, pct))          # order statistics: bit exact


def check_stats(lib, m):
    got, ref = lcl.map_stats(m, lib=lib), ref_map_stats(m)
    assert (got["x_corr"], got["z_corr"]) == (ref["x_corr"], ref["z_corr"])
    for k in ("avg", "rms", "dev"):
        assert got[k] == pytest.approx(ref[k], rel=1e-12, abs=1e-300), k


def cases():
    rng = np.random.default_rng(7)
    yield "random", rng.standard_normal(5000) * 1e-9
    yield "ties", np.round(rng.standard_normal(4000), 1)                  # many duplicates
    yield "signs and zeros", np.concatenate([rng.standard_normal(300), np.zeros(200), -np.zeros(50), [1e300, -1e300, 5e-324]])
    yield "constant", np.full(777, 3.25)
    yield "single", np.array([42.0])
    yield "flux-like", np.abs(rng.standard_normal(3000)) * 1e-14 * (rng.random(3000) > 0.2)


@pytest.fixture(scope="module")
def sim():
    from sim import sim_lib
    return sim_lib.load()


@pytest.mark.parametrize("name,values", list(cases()), ids=[c[0] for c in cases()])
def test_sim_percentiles(sim, name, values):
    check_percentiles(sim, values)
    check_percentiles(sim, values, pct=[33.3, 66.6, 12.5] * 15)          # more prefixes than one pass holds (45 > 32)


@pytest.mark.parametrize("shape", [(40, 56), (33, 17), (1, 64), (64, 1)])
def test_sim_map_stats(sim, shape):
    check_stats(sim, synth_aperture(shape[0], shape[1], seed=3, sigma_c=3) * 1e-5)


def test_sim_map_stats_fixture_and_constant_map(sim):
    check_stats(sim, load_map("Fracture1ApertureMap-10avg.txt"))
    got = lcl.map_stats(np.full((20, 30), 1e-5), lib=sim)
    # a constant map never passes 0.9 of its (zero) variance: the lag runs out at min(nx, nz) / 2 + 1 (SURVEY 8a quirk ix)
    assert (got["x_corr"], got["z_corr"]) == (11.0, 11.0) and got["dev"] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name,values", list(cases()), ids=[c[0] for c in cases()])
def test_gpu_percentiles(name, values):
    check_percentiles(None, values)


@pytest.mark.gpu
def test_gpu_percentiles_and_stats_large():
    import torch
    a = synth_aperture(1500, 1300, seed=11, sigma_c=8, contact=0.2, sigma_k=16)
    check_percentiles(None, a)
    check_stats(None, a * 1e-5)
    t = torch.from_numpy(a).cuda()
    assert np.array_equal(lcl.percentiles(t, PCT), ref_percentiles(a, PCT))        # device pointer in
    assert lcl.map_stats(t)["x_corr"] == ref_map_stats(a)["x_corr"]
