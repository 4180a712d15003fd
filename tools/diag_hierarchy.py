"""Diagnostic (CPU, through the CUDA-on-CPU sim build of the product kernels): pulls the multigrid
hierarchy the product built -- stored 9-point operators and interpolation weights of every level --
out of a solver, reassembles it as scipy matrices and analyses it:
  * Galerkin consistency  A_{l+1} == P^T A_l P
  * PCG iteration count / smallest Ritz values when level k is solved exactly (which level loses modes)
  * the slowest mode of each two-level step
Design evidence only -- not product, not oracle.   usage: python tools/diag_hierarchy.py N CONTACT [SEED]"""
import ctypes as C
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import synth_aperture          # noqa: E402
from apmapflow_b200._lib import Params      # noqa: E402
from apmapflow_b200 import lcl              # noqa: E402
from sim import sim_lib                     # noqa: E402


def field(solver, name):
    ptr, cnt = C.c_void_p(), C.c_size_t()
    rc = solver.lib.apm_lcl_solver_device_field(solver._h, name.encode(), C.byref(ptr), C.byref(cnt))
    if rc != 0:
        raise RuntimeError(name + ": " + (solver.lib.apm_lcl_last_error() or b"").decode())
    return ptr.value, cnt.value


def array(solver, name):
    ptr, cnt = field(solver, name)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(cnt,)).copy()   # sim: host memory


def fine_matrix(tx, tz, txl, tzb, nz, nx):
    """SPD 5-point matrix from faces: tx[z,x] right face of (z,x), txl left boundary, tz up face, tzb bottom."""
    tx = tx.reshape(nz, nx); tz = tz.reshape(nz, nx)
    left = np.concatenate([txl[:, None], tx[:, :-1]], axis=1)
    down = np.concatenate([tzb[None, :], tz[:-1, :]], axis=0)
    diag = (left + tx) + (down + tz)
    idx = np.arange(nz * nx).reshape(nz, nx)
    rows = [idx.ravel(), idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()]
    cols = [idx.ravel(), idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()]
    vals = [diag.ravel(), -tx[:, :-1].ravel(), -tx[:, :-1].ravel(), -tz[:-1, :].ravel(), -tz[:-1, :].ravel()]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nz * nx,) * 2)


def coarse_matrix(Cd, E, N, NE, NW, nz, nx):
    idx = np.arange(nz * nx).reshape(nz, nx)
    E = E.reshape(nz, nx); N = N.reshape(nz, nx); NE = NE.reshape(nz, nx); NW = NW.reshape(nz, nx)
    rows, cols, vals = [idx.ravel()], [idx.ravel()], [Cd]
    for a, b, g in ((idx[:, :-1], idx[:, 1:], E[:, :-1]), (idx[:-1, :], idx[1:, :], N[:-1, :]),
                    (idx[:-1, :-1], idx[1:, 1:], NE[:-1, :-1]), (idx[:-1, 1:], idx[1:, :-1], NW[:-1, 1:])):
        rows += [a.ravel(), b.ravel()]; cols += [b.ravel(), a.ravel()]; vals += [-g.ravel(), -g.ravel()]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nz * nx,) * 2)


def prolongation(w, nz, nx, nzc, nxc):
    cid = np.arange(nzc * nxc).reshape(nzc, nxc)
    fid = np.arange(nz * nx).reshape(nz, nx)
    W = {k: v.reshape(nzc, nxc) for k, v in w.items()}
    rows, cols, vals = [], [], []

    def add(fz, fx, cz, cx, val):   # index arrays over coarse blocks
        ok = (fz < nz) & (fx < nx) & (cz < nzc) & (cx < nxc)
        rows.append(fid[fz[ok], fx[ok]]); cols.append(cid[cz[ok], cx[ok]]); vals.append(val[ok])
    Z, X = np.indices((nzc, nxc))
    Zr, Xr = Z.ravel(), X.ravel()
    one = np.ones(Zr.size)
    add(2 * Zr, 2 * Xr, Zr, Xr, one)
    add(2 * Zr, 2 * Xr + 1, Zr, Xr, W["hw"].ravel()); add(2 * Zr, 2 * Xr + 1, Zr, Xr + 1, W["he"].ravel())
    add(2 * Zr + 1, 2 * Xr, Zr, Xr, W["vs"].ravel()); add(2 * Zr + 1, 2 * Xr, Zr + 1, Xr, W["vn"].ravel())
    add(2 * Zr + 1, 2 * Xr + 1, Zr, Xr, W["csw"].ravel()); add(2 * Zr + 1, 2 * Xr + 1, Zr, Xr + 1, W["cse"].ravel())
    add(2 * Zr + 1, 2 * Xr + 1, Zr + 1, Xr, W["cnw"].ravel()); add(2 * Zr + 1, 2 * Xr + 1, Zr + 1, Xr + 1, W["cne"].ravel())
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nz * nx, nzc * nxc))


class Hierarchy:
    def __init__(self, solver):
        nlev = field(solver, "nlevels")[1]
        self.shape = [(field(solver, "L%d.nz" % l)[1], field(solver, "L%d.nx" % l)[1]) for l in range(nlev)]
        nz, nx = self.shape[0]
        self.A = [fine_matrix(array(solver, "tx"), array(solver, "tz"), array(solver, "txl"), array(solver, "tzb"), nz, nx)]
        self.S = [None]
        for l in range(1, nlev):
            nz, nx = self.shape[l]
            g = {k: array(solver, "L%d.%s" % (l, k)) for k in ("C", "S", "E", "N", "NE", "NW")}
            self.A.append(coarse_matrix(g["C"], g["E"], g["N"], g["NE"], g["NW"], nz, nx))
            self.S.append(g["S"])
        self.P = []
        for l in range(nlev - 1):
            w = {k: array(solver, "L%d.%s" % (l, k)) for k in ("hw", "he", "vs", "vn", "csw", "cse", "cnw", "cne")}
            self.P.append(prolongation(w, *self.shape[l], *self.shape[l + 1]))
        self.colors = []
        for lev, (nz, nx) in enumerate(self.shape):
            zz, xx = np.indices((nz, nx))
            c = ((zz + xx) % 2).ravel() if lev == 0 else ((zz % 2) * 2 + (xx % 2)).ravel()
            self.colors.append([np.nonzero(c == k)[0] for k in range(2 if lev == 0 else 4)])
        self.diag = [a.diagonal() for a in self.A]

    def smooth(self, lev, x, b, forward, nu=1):
        A, d = self.A[lev], self.diag[lev]
        for _ in range(nu):
            for c in (self.colors[lev] if forward else self.colors[lev][::-1]):
                x[c] += (b[c] - A[c] @ x) / d[c]

    def cycle(self, lev, b, exact_level, lu, nu):
        if lev == exact_level:
            return lu.solve(b)
        x = np.zeros_like(b)
        self.smooth(lev, x, b, True, nu if lev else 1)
        r = b - self.A[lev] @ x
        x += self.P[lev] @ self.cycle(lev + 1, self.P[lev].T @ r, exact_level, lu, nu)
        self.smooth(lev, x, b, False, nu if lev else 1)
        return x


def pcg_ritz(A, b, pre, tol=1e-10, maxit=200):
    x = np.zeros_like(b); r = b.copy(); z = pre(r); d = z.copy(); rz = r @ z
    al, be, bn = [], [], np.linalg.norm(b)
    for it in range(maxit):
        q = A @ d; alpha = rz / (d @ q); x += alpha * d; r -= alpha * q; al.append(alpha)
        if np.linalg.norm(r) / bn < tol:
            break
        z = pre(r); rzn = r @ z; beta = rzn / rz; be.append(beta); d = z + beta * d; rz = rzn
    m = len(al); T = np.zeros((m, m))
    for k in range(m):
        T[k, k] = 1 / al[k] + (be[k - 1] / al[k - 1] if k > 0 else 0)
        if k + 1 < m:
            T[k, k + 1] = T[k + 1, k] = np.sqrt(be[k]) / al[k]
    return m, np.sort(np.linalg.eigvalsh(T))


if __name__ == "__main__":
    n = int(sys.argv[1]); contact = float(sys.argv[2]); seed = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    a = synth_aperture(n, n, seed=seed, contact=contact)
    p = Params.defaults(inlet_pressure=1000., outlet_pressure=0., density=1000., viscosity=1e-3, voxel_size=1e-5,
                        avg_factor=1., high_mask=100., low_mask=1e-4, outlet_side='TOP')
    solver = lcl.LclSolver(n, n, lib=sim_lib.load())
    f, s = solver.solve(p, a, fields=("pressure",))
    print("product solve (sim): iterations", s["iterations"], "converged", s["converged"])
    H = Hierarchy(solver)
    print("levels", H.shape)
    for l in range(len(H.P)):
        G = (H.P[l].T @ H.A[l] @ H.P[l]).tocsr()
        D = (G - H.A[l + 1]).tocoo()
        sc = np.sqrt(np.abs(H.diag[l + 1][D.row] * H.diag[l + 1][D.col]))
        rowsum = np.abs(np.asarray(H.A[l + 1].sum(axis=1)).ravel() - H.S[l + 1]) / H.diag[l + 1]
        offpos = (H.A[l + 1] - sp.diags(H.diag[l + 1])).max()
        print("level %d: max |A_c - P^T A P| / sqrt(d_i d_j) = %.2e   max |rowsum - S|/diag = %.2e   largest off-diagonal %.2e"
              % (l + 1, np.abs(D.data / sc).max() if D.nnz else 0.0, rowsum.max(), offpos))
    # right-hand side of the TOP outlet problem
    nz, nx = H.shape[0]
    tzb = array(solver, "tzb")
    b = np.zeros(nz * nx); b[:nx] = tzb * 1000.0
    nu = 2
    for k in range(1, len(H.A)):
        lu = spl.splu(H.A[k].tocsc())
        m, ev = pcg_ritz(H.A[0], b, lambda r: H.cycle(0, r, k, lu, nu))
        print("exact at level %d %-12s iters %3d  smallest ritz: %s" % (k, H.shape[k], m, " ".join("%.3g" % v for v in ev[:5])))
    # ---- slowest mode of a two-level step (level L smoothing + exact level L+1 correction)
    Ld = int(os.environ.get("DIAG_LEVEL", "3"))
    A3, P3 = H.A[Ld], H.P[Ld]
    lu4 = spl.splu(H.A[Ld + 1].tocsc())
    nz3, nx3 = H.shape[Ld]

    def E(e):
        x = e.copy(); bb = np.zeros_like(e)
        H.smooth(Ld, x, bb, True, nu)
        r = bb - A3 @ x; x += P3 @ lu4.solve(P3.T @ r)
        H.smooth(Ld, x, bb, False, nu)
        return x
    rng = np.random.default_rng(0); e = rng.standard_normal(nz3 * nx3)
    for it in range(80):
        e = E(e); e /= np.sqrt(e @ (A3 @ e))
    e2 = E(e)
    print("two-level (L%d, L%d): A-norm factor of the slowest mode %.4f" % (Ld, Ld + 1, np.sqrt(e2 @ (A3 @ e2))))
    y = e.reshape(nz3, nx3); y = y / y.flat[np.argmax(np.abs(y))]
    z0, x0 = np.unravel_index(np.argmax(np.abs(y)), y.shape)
    print("peak at level-%d cell (%d,%d)  [level-%d C-points are the even,even cells]" % (Ld, z0, x0, Ld + 1))
    zs = range(max(z0 - 4, 0), min(z0 + 5, nz3)); xs = range(max(x0 - 5, 0), min(x0 + 6, nx3))
    print("mode y:")
    for zz in zs:
        print("%4d " % zz + " ".join("%7.3f" % y[zz, xx] for xx in xs))
    print("log10 diag:")
    d3 = H.diag[Ld].reshape(nz3, nx3)
    for zz in zs:
        print("%4d " % zz + " ".join("%7.1f" % np.log10(d3[zz, xx]) for xx in xs))
    f = 2 ** Ld
    print("contact fraction of the %dx%d fine block around each cell:" % (f, f))
    for zz in zs:
        print("%4d " % zz + " ".join("%7.2f" % (a[max(zz * f - f // 2, 0):zz * f + f // 2 + 1, max(xx * f - f // 2, 0):xx * f + f // 2 + 1] == 0).mean() for xx in xs))
    for dz, dx in ((0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)):
        if not (0 <= z0 + dz < nz3 and 0 <= x0 + dx < nx3):
            continue
        i = (z0 + dz) * nx3 + x0 + dx
        row = A3.getrow(i).tocoo()
        print("row (%+d,%+d): " % (dz, dx) + "  ".join("(%+d,%+d) %.2e" % (c // nx3 - z0 - dz, c % nx3 - x0 - dx, v) for c, v in zip(row.col, row.data)), " S=%.2e" % H.S[Ld][i])
    # interpolation rows of the F-points around the peak
    for dz in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if not (0 <= z0 + dz < nz3 and 0 <= x0 + dx < nx3):
                continue
            i = (z0 + dz) * nx3 + x0 + dx
            pr = P3.getrow(i).tocoo()
            nxc = H.shape[Ld + 1][1]
            print("P row (%+d,%+d): " % (dz, dx) + "  ".join("C(%d,%d) %.3e" % (c // nxc * 2 - z0, c % nxc * 2 - x0, v) for c, v in zip(pr.col, pr.data)))
    # ---- the same mode in the D^1/2 scaling (where a cell's share of the energy is visible)
    yh = (np.sqrt(H.diag[Ld]) * e).reshape(nz3, nx3)
    yh = yh / np.abs(yh).max()
    zc, xc = np.unravel_index(np.argmax(np.abs(yh)), yh.shape)
    print("D^1/2-scaled mode: peak at (%d,%d); share of |yh|^2 within 5 cells of the peak: %.3f" %
          (zc, xc, (yh[max(zc - 5, 0):zc + 6, max(xc - 5, 0):xc + 6] ** 2).sum() / (yh ** 2).sum()))
    zs = range(max(zc - 5, 0), min(zc + 6, nz3)); xs = range(max(xc - 6, 0), min(xc + 7, nx3))
    for zz in zs:
        print("%4d " % zz + " ".join("%7.3f" % yh[zz, xx] for xx in xs))
    print("log10 diag there:")
    for zz in zs:
        print("%4d " % zz + " ".join("%7.1f" % np.log10(d3[zz, xx]) for xx in xs))
    Ae = (A3 @ e).reshape(nz3, nx3) / np.sqrt(d3)
    print("D^-1/2 A e (scaled residual of the mode), same window, x1e3 of max |yh| scale:")
    sc_ = np.abs(np.sqrt(d3) * e.reshape(nz3, nx3)).max()
    for zz in zs:
        print("%4d " % zz + " ".join("%7.3f" % (Ae[zz, xx] / sc_) for xx in xs))
    if os.environ.get("DIAG_STAGES"):
        print("---- the slow mode through one two-level cycle, D^1/2 scaled, window around the peak")
        sq = np.sqrt(H.diag[Ld])
        def show(tag, v):
            vv = (sq * v).reshape(nz3, nx3) / sc_
            print(tag)
            for zz in zs:
                print("%4d " % zz + " ".join("%7.3f" % vv[zz, xx] for xx in xs))
        x = e.copy(); bb = np.zeros_like(e)
        show("start", x)
        H.smooth(Ld, x, bb, True, nu); show("after pre-smoothing", x)
        r = bb - A3 @ x
        ec = lu4.solve(P3.T @ r)
        corr = P3 @ ec
        show("coarse correction P e_c (to be added)", corr)
        x += corr; show("after coarse correction", x)
        H.smooth(Ld, x, bb, False, nu); show("after post-smoothing", x)
        nzc, nxc = H.shape[Ld + 1]
        print("coarse-level log10 diag around (window/2):")
        d4 = H.diag[Ld + 1].reshape(nzc, nxc)
        for zz in range(zs[0] // 2, zs[-1] // 2 + 1):
            print("%4d " % zz + " ".join("%7.1f" % np.log10(d4[zz, xx]) for xx in range(xs[0] // 2, xs[-1] // 2 + 1)))
