"""Prototype 3 (numpy/scipy, CPU): why MG-PCG stalls below 1e-10 on maps with contact regions, and
the fix.  Open pockets enclosed by LOW-MASK contact cells ("islands") are coupled to the rest by
conductances ~1e-14 of their internal ones; an island without a coarse grid point has a constant
mode that neither the smoother nor the coarse grids see.  Adding one indicator vector per island
(strong-face connected components not anchored to a Dirichlet boundary) as an additive coarse space,
M^-1 = M_mg^-1 + W (W^T A W)^-1 W^T, removes the late CG bumps.  Design evidence only."""
import sys, time
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
import proto_boxmg as pb


def islands(L, theta=1e-4):
    nz, nx = L.nz, L.nx
    n = nz * nx
    idx = np.arange(n).reshape(nz, nx)
    m = np.maximum(np.maximum(L.tl, L.tx), np.maximum(L.td, L.tz))   # max coupling per cell (incl. boundary faces)
    # interior faces
    fx = L.tx[:, :-1]; sx = fx >= theta * np.maximum(m[:, :-1], m[:, 1:])
    fz = L.tz[:-1, :]; sz = fz >= theta * np.maximum(m[:-1, :], m[1:, :])
    rows = np.concatenate([idx[:, :-1][sx], idx[:-1, :][sz]])
    cols = np.concatenate([idx[:, 1:][sx], idx[1:, :][sz]])
    G = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n))
    ncomp, lab = connected_components(G, directed=False)
    lab = lab.reshape(nz, nx)
    # anchored components: strong Dirichlet boundary face
    anchored = np.zeros(ncomp, bool)
    for cells, t in ((lab[0, :], L.tzb), (lab[-1, :], L.tz[-1, :]), (lab[:, 0], L.txl), (lab[:, -1], L.tx[:, -1])):
        mm = {0: m[0, :], 1: m[-1, :], 2: m[:, 0], 3: m[:, -1]}
    anchored[lab[0, :][L.tzb >= theta * m[0, :]]] = True
    anchored[lab[-1, :][L.tz[-1, :] >= theta * m[-1, :]]] = True
    anchored[lab[:, 0][(L.txl >= theta * m[:, 0]) & (L.txl > 0)]] = True
    anchored[lab[:, -1][(L.tx[:, -1] >= theta * m[:, -1]) & (L.tx[:, -1] > 0)]] = True
    size = np.bincount(lab.ravel(), minlength=ncomp)
    keep = (~anchored) & (size >= 2)
    newid = -np.ones(ncomp, int)
    newid[keep] = np.arange(keep.sum())
    return newid[lab], int(keep.sum()), size[keep]


def run(n, contact, seed=0, tol=1e-13, use_w=True, torch_recipe=True):
    if torch_recipe:
        import bench
        a = bench.synth_map_torch(n, seed, "cpu", contact=contact).numpy()
    else:
        a = pb.synth_map(n, seed=seed, contact=contact)
    tx, tz, txl, tzb, w = pb.conductances(a)
    L = pb.Level(tx, tz, txl, tzb)
    A = pb.level_matrix(L)
    b = pb.rhs_for(L, "TOP", 1000.0, 0.0).ravel()
    dinv = 1.0 / A.diagonal()
    mg = pb.GMG(A, n, n, kind="box", nu=1, coarsest=8)
    lab, k, sizes = islands(L)
    print("n=%d islands=%d (cells in islands %d, largest %d)" % (n, k, sizes.sum(), sizes.max() if k else 0))
    labf = lab.ravel()
    sel = labf >= 0
    W = sp.csr_matrix((np.ones(sel.sum()), (np.nonzero(sel)[0], labf[sel])), shape=(n * n, max(k, 1)))
    E = (W.T @ A @ W).diagonal()
    E[E == 0] = 1.0

    def precond(r):
        z = mg.vcycle(0, r)
        if use_w and k:
            z = z + W @ ((W.T @ r) / E)
        return z
    x = np.zeros_like(b); r = b.copy(); z = precond(r); d = z.copy(); rz = r @ z
    bn = np.linalg.norm(b); hist = []
    for it in range(1, 200):
        q = A @ d; alpha = rz / (d @ q); x += alpha * d; r -= alpha * q
        rn = np.linalg.norm(r) / bn; dx = np.abs(alpha * d).max()
        hist.append((rn, dx))
        if rn < tol: break
        z = precond(r); rz_new = r @ z; d = z + (rz_new / rz) * d; rz = rz_new
    print("  use_w=%s iters=%d" % (use_w, it), " ".join("%.0e/%.0e" % h for h in hist))
    return x


if __name__ == "__main__":
    n = int(sys.argv[1]); c = float(sys.argv[2])
    x0 = run(n, c, use_w=False)
    x1 = run(n, c, use_w=True)
    print("max diff", np.abs(x0 - x1).max())
