"""Prototype 2 (numpy/scipy, CPU): operator-dependent (BoxMG-style) interpolation with
Galerkin 9-point coarse operators, compared with plain 2x2 aggregation (proto_mg.py).
Design evidence only -- not product, not oracle."""
import sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, '/root/repo/tools')
from proto_mg import synth_map, conductances, Level, rhs_for, direct


def level_matrix(L):
    n = L.nz * L.nx
    idx = np.arange(n).reshape(L.nz, L.nx)
    rows = [idx.ravel(), idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()]
    cols = [idx.ravel(), idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()]
    vals = [L.diag.ravel(), -L.tl[:, 1:].ravel(), -L.tx[:, :-1].ravel(), -L.td[1:, :].ravel(), -L.tz[:-1, :].ravel()]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def stencil_arrays(A, nz, nx):
    """return dict (dz,dx)->array[nz,nx] of coefficient from (z,x) to (z+dz,x+dx) (0 if absent)"""
    A = A.tocoo()
    zi, xi = np.divmod(A.row, nx)
    zj, xj = np.divmod(A.col, nx)
    st = {}
    for dz in (-1, 0, 1):
        for dx in (-1, 0, 1):
            m = (zj - zi == dz) & (xj - xi == dx)
            arr = np.zeros((nz, nx))
            arr[zi[m], xi[m]] = A.data[m]
            st[(dz, dx)] = arr
    return st


def boxmg_P(A, nz, nx):
    st = stencil_arrays(A, nz, nx)
    nzc, nxc = (nz + 1) // 2, (nx + 1) // 2
    cidx = lambda Z, X: Z * nxc + X
    rows, cols, vals = [], [], []
    W = np.zeros((nz, nx, 3, 3))  # weight of fine (z,x) from coarse neighbour offset (dz,dx) in fine units /1
    # helper returning weight arrays for coarse points located at fine offsets
    zz, xx = np.indices((nz, nx))
    ez, ex = (zz % 2 == 0), (xx % 2 == 0)
    # coarse points
    w = {}
    for dz in (-1, 0, 1):
        for dx in (-1, 0, 1):
            w[(dz, dx)] = np.zeros((nz, nx))
    w[(0, 0)][ez & ex] = 1.0
    # h-edge points: z even, x odd
    m = ez & ~ex
    den = st[(0, 0)] + st[(1, 0)] + st[(-1, 0)]
    wl = -(st[(-1, -1)] + st[(0, -1)] + st[(1, -1)]) / den
    wr = -(st[(-1, 1)] + st[(0, 1)] + st[(1, 1)]) / den
    w[(0, -1)][m] = wl[m]
    w[(0, 1)][m] = wr[m]
    # v-edge points: z odd, x even
    m = ~ez & ex
    den = st[(0, 0)] + st[(0, 1)] + st[(0, -1)]
    wd = -(st[(-1, -1)] + st[(-1, 0)] + st[(-1, 1)]) / den
    wu = -(st[(1, -1)] + st[(1, 0)] + st[(1, 1)]) / den
    w[(-1, 0)][m] = wd[m]
    w[(1, 0)][m] = wu[m]
    # centre points: z odd, x odd: e = -sum a_nbr e_nbr / a_C with edge nbrs already interpolated
    m = ~ez & ~ex
    aC = st[(0, 0)]

    def sh(arr, dz, dx):  # arr shifted so result[z,x] = arr[z+dz, x+dx] (0 outside)
        out = np.zeros_like(arr)
        zs = slice(max(0, -dz), nz - max(0, dz)); zt = slice(max(0, dz), nz - max(0, -dz))
        xs = slice(max(0, -dx), nx - max(0, dx)); xt = slice(max(0, dx), nx - max(0, -dx))
        out[zs, xs] = arr[zt, xt]
        return out
    for (cz, cx) in [(-1, -1), (-1, 1), (1, -1), (1, 1)]:
        # direct corner + via horizontal-edge nbr (z+cz, x) whose coarse nbr in x dir cx, + via vertical-edge nbr (z, x+cx)
        acc = st[(cz, cx)].copy()
        acc += st[(cz, 0)] * sh(w[(0, cx)], cz, 0)   # nbr (z+cz, x) is an h-edge point (z+cz even, x odd)
        acc += st[(0, cx)] * sh(w[(cz, 0)], 0, cx)   # nbr (z, x+cx) is a v-edge point
        w[(cz, cx)][m] = (-acc / aC)[m]
    for (dz, dx), arr in w.items():
        zc, xc = zz + dz, xx + dx
        ok = (arr != 0) & (zc >= 0) & (zc < nz) & (xc >= 0) & (xc < nx)
        ok &= (zc % 2 == 0) & (xc % 2 == 0)
        rows.append((zz * nx + xx)[ok])
        cols.append(cidx(zc[ok] // 2, xc[ok] // 2))
        vals.append(arr[ok])
    P = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nz * nx, nzc * nxc))
    return P, nzc, nxc


def agg_P(nz, nx):
    nzc, nxc = (nz + 1) // 2, (nx + 1) // 2
    zz, xx = np.indices((nz, nx))
    rows = (zz * nx + xx).ravel()
    cols = ((zz // 2) * nxc + xx // 2).ravel()
    return sp.csr_matrix((np.ones(nz * nx), (rows, cols)), shape=(nz * nx, nzc * nxc)), nzc, nxc


class GMG:
    def __init__(self, A, nz, nx, kind="box", coarsest=8, nu=1, smoother="gs", omega=0.8, scale=2.0, fine_only_box=False):
        self.A, self.P, self.shape, self.diag, self.colors = [A], [], [(nz, nx)], [A.diagonal()], []
        self.nu, self.smoother, self.omega = nu, smoother, omega
        while max(self.shape[-1]) > coarsest:
            nz, nx = self.shape[-1]
            if kind == "box":
                P, nzc, nxc = boxmg_P(self.A[-1], nz, nx)
                Ac = (P.T @ self.A[-1] @ P).tocsr()
            else:
                P, nzc, nxc = agg_P(nz, nx)
                Ac = (P.T @ self.A[-1] @ P).tocsr() / scale
            Ac.eliminate_zeros()
            self.P.append(P); self.A.append(Ac); self.shape.append((nzc, nxc)); self.diag.append(Ac.diagonal())
        for lev, (nz, nx) in enumerate(self.shape):
            zz, xx = np.indices((nz, nx))
            if lev == 0 or kind != "box":
                c = ((zz + xx) % 2).ravel(); nc = 2
            else:
                c = ((zz % 2) * 2 + (xx % 2)).ravel(); nc = 4
            self.colors.append([np.nonzero(c == k)[0] for k in range(nc)])
        self.Ainv = np.linalg.pinv(self.A[-1].toarray())
        self.nnz = [a.nnz for a in self.A]

    def smooth(self, lev, x, b, forward):
        A, d = self.A[lev], self.diag[lev]
        for _ in range(self.nu):
            if self.smoother == "jac":
                x += self.omega * (b - A @ x) / d
            else:
                cols = self.colors[lev] if forward else self.colors[lev][::-1]
                for c in cols:
                    Ac = A[c]
                    x[c] += (b[c] - Ac @ x) / d[c]

    def vcycle(self, lev, b):
        if lev == len(self.A) - 1:
            return self.Ainv @ b
        x = np.zeros_like(b)
        self.smooth(lev, x, b, True)
        r = b - self.A[lev] @ x
        ec = self.vcycle(lev + 1, self.P[lev].T @ r)
        x += self.P[lev] @ ec
        self.smooth(lev, x, b, False)
        return x


def pcg(A, b, precond, dinv, tol=1e-10, maxit=3000, x_true=None):
    x = np.zeros_like(b); r = b.copy(); z = precond(r); d = z.copy(); rz = r @ z
    bn = np.linalg.norm(b); bs = np.sqrt((b * b * dinv).sum())
    for it in range(1, maxit + 1):
        q = A @ d
        alpha = rz / (d @ q)
        x += alpha * d; r -= alpha * q
        rn = np.linalg.norm(r) / bn; rs = np.sqrt((r * r * dinv).sum()) / bs
        if rn < tol and rs < tol:
            break
        z = precond(r); rz_new = r @ z; d = z + (rz_new / rz) * d; rz = rz_new
    err = None if x_true is None else np.abs(x - x_true).max() / np.abs(x_true).max()
    return x, it, (rn, rs, err)


if __name__ == "__main__":
    n = int(sys.argv[1]); contact = float(sys.argv[2])
    a = synth_map(n, contact=contact)
    tx, tz, txl, tzb, w = conductances(a)
    L = Level(tx, tz, txl, tzb)
    A = level_matrix(L)
    b = rhs_for(L, "TOP", 1000.0, 0.0).ravel()
    dinv = 1.0 / A.diagonal()
    for name, kw in [("box gs nu1", dict(kind="box", nu=1)),
                     ("box jac nu2", dict(kind="box", nu=2, smoother="jac")),
                     ("agg gs nu1", dict(kind="agg", nu=1))]:
        t0 = time.time()
        mg = GMG(A, n, n, **kw)
        t1 = time.time()
        x, it, fin = pcg(A, b, lambda r: mg.vcycle(0, r), dinv)
        print(f"n={n} c={contact} {name}: levels {len(mg.A)} iters {it} final {fin} setup {t1-t0:.1f}s solve {time.time()-t1:.1f}s", flush=True)
