"""Numerical prototype (numpy, CPU) used to choose the multigrid components before
any CUDA was written.  NOT part of the product path and NOT the oracle: it only
answers "which smoother / coarse operator / over-correction keeps MG-PCG iteration
counts flat under 1e14+ conductance contrast".  Results are summarised in DESIGN.md.

Layout used here is the one the CUDA kernels use:
    tx[z, x]  = conductance of the RIGHT face of cell (z, x)   (x = nx-1 -> right boundary face)
    tz[z, x]  = conductance of the TOP   face of cell (z, x)   (z = nz-1 -> top boundary face)
    txl[z]    = left boundary face of row z,  tzb[x] = bottom boundary face of column x
"""
import sys
import time
import numpy as np


# ----------------------------------------------------------------------------
# synthetic maps (SURVEY.md 8d recipe, numpy flavour)
def gauss_smooth(z, sigma):
    if sigma <= 0:
        return z
    r = int(4 * sigma)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    from scipy.ndimage import convolve1d
    z = convolve1d(z, k, axis=0, mode="wrap")
    z = convolve1d(z, k, axis=1, mode="wrap")
    return (z - z.mean()) / z.std()


def synth_map(n, seed=0, sigma_c=8, contact=0.2, sigma_k=16):
    rng = np.random.default_rng(20240601 + seed)
    z = gauss_smooth(rng.standard_normal((n, n)), sigma_c)
    a = np.exp(np.log(5.0) + 0.5 * z)
    a = np.minimum(a, 100.0)
    if contact > 0:
        c = gauss_smooth(rng.standard_normal((n, n)), sigma_k)
        thr = np.quantile(c, contact)
        a[c < thr] = 0.0
    return a


# ----------------------------------------------------------------------------
# conductances (vectorised restatement of READ_AP_MAP/CREATE_POINT_MAP/COEF/FLOW_BOUNDARY)
def conductances(a, vox=1e-5, avg_fact=1.0, rough=0.0, low=1e-4, high=100.0, visc=1e-3, side="TOP"):
    nz, nx = a.shape
    a = np.where(a > high, high, a)
    a = np.where(a < low, low, a)
    # corner map (nz+1, nx+1)
    ap = np.pad(a, 1, mode="edge")
    # interior corners: 4-avg ; edges: the pad replicates so 2-avg == (x+x+y+y)/4 -- not bit exact, fine for proto
    c = (((ap[:-1, :-1] + ap[1:, :-1]) + ap[:-1, 1:]) + ap[1:, 1:]) / 4.0
    w = a - rough
    w = np.where(w < low, low, w) * vox
    rht = 0.5 * (c[:-1, 1:] + c[1:, 1:]) - rough   # BRC + TRC
    top = 0.5 * (c[1:, :-1] + c[1:, 1:]) - rough   # TLC + TRC
    rht = np.where(rht < low, low, rht) * vox
    top = np.where(top < low, low, top) * vox
    d = vox * avg_fact
    dbp = d / 2

    def half(wc, b):
        th = np.arctan(np.abs(wc - b) / dbp)
        w3 = (2 * wc ** 2 * b ** 2) / (wc + b)
        t = np.tan(th)
        x = t - th
        with np.errstate(divide="ignore", invalid="ignore"):
            b3 = w3 * 3 * x / t ** 3
        return np.where(np.abs(1e9 * x) < 1.0, w3, b3)

    tx = np.zeros((nz, nx))
    tz = np.zeros((nz, nx))
    bp = half(w, rht)
    bf = half(w[:, 1:], rht[:, :-1])
    tx[:, :-1] = 1.0 / (dbp / bp[:, :-1] + dbp / bf) * (d / 12)
    tx[:, -1] = d * bp[:, -1] / (12 * d)
    bp = half(w, top)
    bf = half(w[1:, :], top[:-1, :])
    tz[:-1, :] = 1.0 / (dbp / bp[:-1, :] + dbp / bf) * (d / 12)
    tz[-1, :] = d * bp[-1, :] / (12 * d)
    tx /= visc
    tz /= visc
    if side in ("TOP", "BOTTOM"):
        txl = np.zeros(nz)
        tx[:, -1] = 0.0
        tzb = 2 * tz[0, :]
        tz[-1, :] = 2 * tz[-1, :]
    else:
        txl = 2 * tx[:, 0]
        tx[:, -1] = 2 * tx[:, -1]
        tzb = np.zeros(nx)
        tz[-1, :] = 0.0
    return tx, tz, txl, tzb, w


# ----------------------------------------------------------------------------
class Level:
    def __init__(self, tx, tz, txl, tzb):
        self.tx, self.tz, self.txl, self.tzb = tx, tz, txl, tzb
        self.nz, self.nx = tx.shape
        # face conductances seen from each cell
        self.tl = np.concatenate([txl[:, None], tx[:, :-1]], axis=1)
        self.td = np.concatenate([tzb[None, :], tz[:-1, :]], axis=0)
        self.diag = self.tl + tx + self.td + tz
        zz, xx = np.indices(tx.shape)
        self.red = ((zz + xx) % 2) == 0
        self.black = ~self.red

    def offdiag(self, x):
        s = np.zeros_like(x)
        s[:, 1:] += self.tl[:, 1:] * x[:, :-1]
        s[:, :-1] += self.tx[:, :-1] * x[:, 1:]
        s[1:, :] += self.td[1:, :] * x[:-1, :]
        s[:-1, :] += self.tz[:-1, :] * x[1:, :]
        return s

    def apply(self, x):
        return self.diag * x - self.offdiag(x)

    def gs_color(self, x, b, mask):
        s = self.offdiag(x)
        x[mask] = (b[mask] + s[mask]) / self.diag[mask]

    def jacobi(self, x, b, omega):
        x += omega * (b - self.apply(x)) / self.diag


def coarsen(L, scale):
    nz, nx = L.nz, L.nx
    nzc, nxc = (nz + 1) // 2, (nx + 1) // 2

    def sum_rows(v):  # (nz, k) -> (nzc, k)
        if nz % 2:
            v = np.concatenate([v, np.zeros((1,) + v.shape[1:])], axis=0)
        return v[0::2] + v[1::2]

    def sum_cols(v):
        if nx % 2:
            v = np.concatenate([v, np.zeros(v.shape[:1] + (1,))], axis=1)
        return v[:, 0::2] + v[:, 1::2]

    # right face of coarse col X = right face of fine col min(2X+1, nx-1)
    cols = np.minimum(2 * np.arange(nxc) + 1, nx - 1)
    rows = np.minimum(2 * np.arange(nzc) + 1, nz - 1)
    txc = sum_rows(L.tx[:, cols]) / scale
    tzc = sum_cols(L.tz[rows, :]) / scale
    txlc = sum_rows(L.txl[:, None])[:, 0] / scale
    tzbc = sum_cols(L.tzb[None, :])[0] / scale
    return Level(txc, tzc, txlc, tzbc)


def restrict(r, nzc, nxc):
    nz, nx = r.shape
    rp = np.zeros((2 * nzc, 2 * nxc))
    rp[:nz, :nx] = r
    return rp[0::2, 0::2] + rp[0::2, 1::2] + rp[1::2, 0::2] + rp[1::2, 1::2]


def prolong(e, nz, nx):
    return np.repeat(np.repeat(e, 2, axis=0), 2, axis=1)[:nz, :nx]


class MG:
    def __init__(self, L0, scale=2.0, nu=1, smoother="rbgs", coarsest=8, omega=0.8, ncoarse_sweeps=30,
                 cycle="V"):
        self.levels = [L0]
        while max(self.levels[-1].nz, self.levels[-1].nx) > coarsest:
            self.levels.append(coarsen(self.levels[-1], scale))
        self.nu, self.smoother, self.omega = nu, smoother, omega
        self.ncs = ncoarse_sweeps
        self.cycle = cycle
        # dense coarsest
        Lc = self.levels[-1]
        n = Lc.nz * Lc.nx
        A = np.zeros((n, n))
        for i in range(n):
            e = np.zeros(n)
            e[i] = 1
            A[:, i] = Lc.apply(e.reshape(Lc.nz, Lc.nx)).ravel()
        self.Ainv = np.linalg.pinv(A)

    def smooth(self, L, x, b, forward, first_zero=False):
        for _ in range(self.nu):
            if self.smoother == "rbgs":
                if forward:
                    L.gs_color(x, b, L.red)
                    L.gs_color(x, b, L.black)
                else:
                    L.gs_color(x, b, L.black)
                    L.gs_color(x, b, L.red)
            else:
                L.jacobi(x, b, self.omega)

    def vcycle(self, lev, b):
        L = self.levels[lev]
        if lev == len(self.levels) - 1:
            return (self.Ainv @ b.ravel()).reshape(b.shape)
        x = np.zeros_like(b)
        self.smooth(L, x, b, True)
        r = b - L.apply(x)
        Lc = self.levels[lev + 1]
        rc = restrict(r, Lc.nz, Lc.nx)
        ec = self.vcycle(lev + 1, rc)
        if self.cycle == "W" and lev + 1 < len(self.levels) - 1:
            # second coarse visit (not symmetric-safe in general; experiment only)
            rc2 = rc - Lc.apply(ec)
            ec = ec + self.vcycle(lev + 1, rc2)
        x += prolong(ec, L.nz, L.nx)
        self.smooth(L, x, b, False)
        return x


def pcg(L, b, precond, tol=1e-10, maxit=2000, x_true=None, verbose=False):
    x = np.zeros_like(b)
    r = b.copy()
    z = precond(r)
    d = z.copy()
    rz = (r * z).sum()
    bn = np.sqrt((b * b).sum())
    dinv = 1.0 / L.diag
    bs = np.sqrt((b * b * dinv).sum())
    hist = []
    for it in range(1, maxit + 1):
        q = L.apply(d)
        alpha = rz / (d * q).sum()
        x += alpha * d
        r -= alpha * q
        rn = np.sqrt((r * r).sum()) / bn
        rs = np.sqrt((r * r * dinv).sum()) / bs
        err = None
        if x_true is not None:
            err = np.abs(x - x_true).max() / np.abs(x_true).max()
        hist.append((rn, rs, err))
        if verbose:
            print(it, rn, rs, err)
        if rn < tol and rs < tol:
            break
        z = precond(r)
        rz_new = (r * z).sum()
        beta = rz_new / rz
        rz = rz_new
        d = z + beta * d
    return x, it, hist


def rhs_for(L, side, p_in, p_out):
    b = np.zeros((L.nz, L.nx))
    if side == "TOP":
        b[0, :] += L.tzb * p_in
        b[-1, :] += L.tz[-1, :] * p_out
    elif side == "BOTTOM":
        b[0, :] += L.tzb * p_out
        b[-1, :] += L.tz[-1, :] * p_in
    elif side == "LEFT":
        b[:, 0] += L.txl * p_out
        b[:, -1] += L.tx[:, -1] * p_in
    else:
        b[:, 0] += L.txl * p_in
        b[:, -1] += L.tx[:, -1] * p_out
    return b


def direct(L, b):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    n = L.nz * L.nx
    idx = np.arange(n).reshape(L.nz, L.nx)
    rows = [idx.ravel()]
    cols = [idx.ravel()]
    vals = [L.diag.ravel()]
    rows += [idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()]
    cols += [idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()]
    vals += [-L.tl[:, 1:].ravel(), -L.tx[:, :-1].ravel(), -L.td[1:, :].ravel(), -L.tz[:-1, :].ravel()]
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    # scale to avoid underflow trouble inside SuperLU
    s = 1.0 / np.sqrt(L.diag.ravel())
    As = sp.diags(s) @ A @ sp.diags(s)
    y = spl.spsolve(As.tocsc(), b.ravel() * s)
    return (y * s).reshape(L.nz, L.nx)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    contact = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    a = synth_map(n, contact=contact)
    tx, tz, txl, tzb, w = conductances(a)
    L = Level(tx, tz, txl, tzb)
    print("n", n, "T range", tx[tx > 0].min(), tx.max())
    b = rhs_for(L, "TOP", 1000.0, 0.0)
    xt = direct(L, b) if n <= 1024 else None
    for (name, kw) in [
        ("rbgs nu1 s2.0", dict(scale=2.0, nu=1)),
        ("rbgs nu1 s1.5", dict(scale=1.5, nu=1)),
        ("rbgs nu1 s1.0", dict(scale=1.0, nu=1)),
        ("rbgs nu2 s2.0", dict(scale=2.0, nu=2)),
        ("jac  nu2 s2.0", dict(scale=2.0, nu=2, smoother="jac", omega=0.8)),
    ]:
        t0 = time.time()
        mg = MG(L, **kw)
        x, it, hist = pcg(L, b, lambda r: mg.vcycle(0, r), x_true=xt)
        print(f"{name}: levels {len(mg.levels)} iters {it}  final {hist[-1]}  {time.time()-t0:.1f}s")
    dinv = 1.0 / L.diag
    x, it, hist = pcg(L, b, lambda r: r * dinv, x_true=xt, maxit=20000)
    print("jacobi-pcg iters", it, hist[-1])
