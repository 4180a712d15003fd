"""Prototype 4 (numpy/scipy, CPU): pattern-restricted relaxation of the BoxMG interpolation ("energy
minimising" within the 3x3 footprint): P <- P - w D^-1 A P on the F rows, entries outside the original
sparsity pattern dropped, row sums restored.  Question: does it remove the outlier modes contact maps
produce?  Design evidence only -- not product, not oracle."""
import sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo/tests')
from proto_mg import conductances, Level, rhs_for
from proto_boxmg import GMG, level_matrix, boxmg_P, pcg
from helpers import synth_aperture


def relax_P(A, P, nz, nx, sweeps=1, omega=1.0):
    zz, xx = np.indices((nz, nx))
    cpt = ((zz % 2 == 0) & (xx % 2 == 0)).ravel()
    pattern = P.copy(); pattern.data[:] = 1.0
    rowsum = np.asarray(P.sum(axis=1)).ravel()
    dinv = 1.0 / A.diagonal()
    for _ in range(sweeps):
        AP = (A @ P).tocsr()
        Pn = (P - sp.diags(omega * dinv) @ AP).multiply(pattern).tocsr()     # keep the pattern only
        rs = np.asarray(Pn.sum(axis=1)).ravel()
        scale = np.where(np.abs(rs) > 1e-300, rowsum / np.where(rs == 0, 1, rs), 1.0)
        Pn = sp.diags(scale) @ Pn
        # C-point rows stay the identity
        keepC = sp.diags(cpt.astype(float)); keepF = sp.diags((~cpt).astype(float))
        P = (keepC @ P + keepF @ Pn).tocsr()
    return P


class RGMG(GMG):
    def __init__(self, A, nz, nx, sweeps=0, omega=1.0, coarsest=8, nu=1, relax_levels=None):
        self.A, self.P, self.shape, self.diag, self.colors = [A], [], [(nz, nx)], [A.diagonal()], []
        self.nu, self.smoother, self.omega = nu, "gs", 0.8
        lev = 0
        while max(self.shape[-1]) > coarsest:
            nz, nx = self.shape[-1]
            P, nzc, nxc = boxmg_P(self.A[-1], nz, nx)
            if sweeps and (relax_levels is None or lev in relax_levels):
                P = relax_P(self.A[-1], P, nz, nx, sweeps, omega)
            Ac = (P.T @ self.A[-1] @ P).tocsr(); Ac.eliminate_zeros()
            self.P.append(P); self.A.append(Ac); self.shape.append((nzc, nxc)); self.diag.append(Ac.diagonal())
            lev += 1
        for lev, (nz, nx) in enumerate(self.shape):
            zz, xx = np.indices((nz, nx))
            c = ((zz + xx) % 2).ravel() if lev == 0 else ((zz % 2) * 2 + (xx % 2)).ravel()
            self.colors.append([np.nonzero(c == k)[0] for k in range(2 if lev == 0 else 4)])
        self.Ainv = np.linalg.pinv(self.A[-1].toarray())


if __name__ == "__main__":
    n = int(sys.argv[1]); contact = float(sys.argv[2]); seed = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    a = synth_aperture(n, n, seed=seed, contact=contact)
    tx, tz, txl, tzb, w = conductances(a)
    L = Level(tx, tz, txl, tzb); A = level_matrix(L)
    b = rhs_for(L, "TOP", 1000.0, 0.0).ravel()
    dinv = 1.0 / A.diagonal()
    for name, kw in (("BoxMG", dict(sweeps=0)), ("BoxMG + 1 relax w=1", dict(sweeps=1, omega=1.0)),
                     ("BoxMG + 1 relax w=0.7", dict(sweeps=1, omega=0.7)), ("BoxMG + 2 relax w=0.7", dict(sweeps=2, omega=0.7))):
        t0 = time.time()
        mg = RGMG(A, n, n, **kw)
        x, it, fin = pcg(A, b, lambda r: mg.vcycle(0, r), dinv)
        print("n=%d c=%.2f seed %d  %-24s iters %3d  (%.0f s)" % (n, contact, seed, name, it, time.time() - t0), flush=True)
