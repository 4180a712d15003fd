"""Prototype 3 (numpy/scipy, CPU): cycle shape (V / W / truncated W / F) against PCG iteration count on
contact maps.  Design evidence only -- not product, not oracle."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo/tools')
sys.path.insert(0, '/root/repo/tests')
from proto_mg import conductances, Level, rhs_for
from proto_boxmg import GMG, level_matrix, pcg
from helpers import synth_aperture


class Cyc(GMG):
    def cycle(self, lev, b, x=None, gamma=lambda l: 1, kind="V"):
        if lev == len(self.A) - 1:
            return self.Ainv @ b
        if x is None:
            x = np.zeros_like(b)
        self.smooth(lev, x, b, True)
        r = b - self.A[lev] @ x
        bc = self.P[lev].T @ r
        g = gamma(lev + 1)
        ec = self.cycle(lev + 1, bc, None, gamma, kind)
        for _ in range(g - 1):
            ec = self.cycle(lev + 1, bc, ec, gamma, kind)
        x += self.P[lev] @ ec
        self.smooth(lev, x, b, False)
        return x


if __name__ == "__main__":
    n = int(sys.argv[1]); contact = float(sys.argv[2])
    a = synth_aperture(n, n, seed=7, contact=contact)
    tx, tz, txl, tzb, w = conductances(a)
    L = Level(tx, tz, txl, tzb)
    A = level_matrix(L)
    b = rhs_for(L, "TOP", 1000.0, 0.0).ravel()
    dinv = 1.0 / A.diagonal()
    mg = Cyc(A, n, n, kind="box", nu=1, coarsest=8)
    nl = len(mg.A)
    print("levels", nl, [s for s in mg.shape], flush=True)
    variants = [("V", lambda l: 1)]
    import os
    if os.environ.get("PROTO_SET", "a") == "a":
        for lo in (1, 2, 3):
            variants.append((f"W from L{lo}", lambda l, lo=lo: 2 if l >= lo else 1))
        for lo, hi in ((2, 4), (2, 5), (3, 5), (3, 6), (2, 6)):
            variants.append((f"W L{lo}..L{hi}", lambda l, lo=lo, hi=hi: 2 if lo <= l <= hi else 1))
    else:
        for lv in (1, 2, 3, 4):
            variants.append((f"gamma2 at L{lv} only", lambda l, lv=lv: 2 if l == lv else 1))
        for lo, hi in ((1, 2), (2, 3), (3, 4), (1, 3)):
            variants.append((f"W L{lo}..L{hi}", lambda l, lo=lo, hi=hi: 2 if lo <= l <= hi else 1))
        variants.append(("gamma3 at L2 only", lambda l: 3 if l == 2 else 1))
    for name, g in variants:
        t1 = time.time()
        x, it, fin = pcg(A, b, lambda r: mg.cycle(0, r, None, g), dinv)
        print(f"n={n} c={contact} {name}: iters {it} rn {fin[0]:.2e} solve {time.time()-t1:.1f}s", flush=True)
