"""Prototype 4 (numpy/scipy): cancellation-free ("difference form") multigrid hierarchy.
Every level keeps positive couplings g and an absorption term S with diag = S + sum(g); operators
are applied as sum g (x_p - x_q) + S x_p and coarse diagonals are rebuilt from coarse couplings and
a separately restricted absorption, so leak conductances 1e-14 below the internal ones of a pocket
are not lost to rounding.  Design evidence only."""
import sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
import proto_boxmg as pb


class DLevel:
    def __init__(self, A_off, S, shape):
        # A_off: csr with NEGATIVE off-diagonals only (zero diagonal); S absorption >= 0
        self.shape = shape
        coo = A_off.tocoo()
        self.rows, self.cols, self.g = coo.row, coo.col, -coo.data
        self.S = S
        self.n = S.size
        self.diag = S + np.bincount(self.rows, self.g, minlength=self.n)
        self.G = sp.csr_matrix((self.g, (self.rows, self.cols)), shape=(self.n, self.n))
        nz, nx = shape
        zz, xx = np.indices(shape)
        self.colors = [np.nonzero(((zz % 2) * 2 + (xx % 2)).ravel() == k)[0] for k in range(4)]

    def csr(self):
        return (sp.diags(self.diag) - self.G).tocsr()

    def matvec(self, x):
        return self.S * x + np.bincount(self.rows, self.g * (x[self.rows] - x[self.cols]), minlength=self.n)

    def gs(self, x, b, forward):
        for c in (self.colors if forward else self.colors[::-1]):
            x[c] = (b[c] + self.G[c] @ x) / self.diag[c]


def build(L, coarsest=8, snap=1e-12, fix=True):
    A = pb.level_matrix(L)
    n = L.nz * L.nx
    S = np.zeros((L.nz, L.nx))
    S[0, :] += L.tzb; S[-1, :] += L.tz[-1, :]; S[:, 0] += L.txl; S[:, -1] += L.tx[:, -1]
    off = A - sp.diags(A.diagonal())
    levels = [DLevel(off.tocsr(), S.ravel(), (L.nz, L.nx))]
    Ps = []
    while max(levels[-1].shape) > coarsest:
        lv = levels[-1]
        Acsr = lv.csr()
        P, nzc, nxc = pb.boxmg_P(Acsr, *lv.shape)
        Ac = (P.T @ Acsr @ P).tocsr()
        offc = Ac - sp.diags(Ac.diagonal())
        offc.eliminate_zeros()
        if fix:
            t = 1.0 - np.asarray(P.sum(axis=1)).ravel()
            t[np.abs(t) < snap] = 0.0
            s = 1.0 - t
            As = lv.matvec(s)
            Sc = P.T @ As
            Sc[np.abs(Sc) < 1e-300] = 0.0
        else:
            Sc = np.asarray(Ac.sum(axis=1)).ravel()
        levels.append(DLevel(offc.tocsr(), np.maximum(Sc, 0.0) if fix else Sc, (nzc, nxc)))
        Ps.append(P)
    Ainv = np.linalg.pinv(levels[-1].csr().toarray())
    return levels, Ps, Ainv


def vcycle(levels, Ps, Ainv, lev, b, nu=1):
    lv = levels[lev]
    if lev == len(levels) - 1:
        return Ainv @ b
    x = np.zeros_like(b)
    for _ in range(nu): lv.gs(x, b, True)
    r = b - lv.matvec(x)
    x += Ps[lev] @ vcycle(levels, Ps, Ainv, lev + 1, Ps[lev].T @ r, nu)
    for _ in range(nu): lv.gs(x, b, False)
    return x


def run(n, contact, fix, tol=1e-14, low=1e-4, seed=0):
    import bench
    a = bench.synth_map_torch(n, seed, "cpu", contact=contact).numpy()
    tx, tz, txl, tzb, w = pb.conductances(a, low=low)
    L = pb.Level(tx, tz, txl, tzb)
    b = pb.rhs_for(L, "TOP", 1000.0, 0.0).ravel()
    t0 = time.time()
    levels, Ps, Ainv = build(L, fix=fix)
    f = levels[0]
    x = np.zeros_like(b); r = b.copy(); z = vcycle(levels, Ps, Ainv, 0, r); d = z.copy(); rz = r @ z
    bn = np.linalg.norm(b); h = []
    for it in range(1, 120):
        q = f.matvec(d); alpha = rz / (d @ q); x += alpha * d; r -= alpha * q
        rn = np.linalg.norm(r) / bn; h.append(rn)
        if rn < tol: break
        z = vcycle(levels, Ps, Ainv, 0, r); rzn = r @ z; d = z + (rzn / rz) * d; rz = rzn
    print('n', n, 'fix', fix, 'iters', it, ' '.join('%.0e' % v for v in h), '%.0fs' % (time.time() - t0), flush=True)
    return x


if __name__ == "__main__":
    n = int(sys.argv[1])
    x0 = run(n, 0.2, False)
    x1 = run(n, 0.2, True)
    print('diff', np.abs(x0 - x1).max())
