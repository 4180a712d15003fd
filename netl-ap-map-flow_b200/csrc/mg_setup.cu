// mg_setup.cu -- builds the multigrid hierarchy for one set of face conductances.
//
// New work (the reference solves the system directly, D4_SOLVER_MODULE.F:444-617): a geometric
// hierarchy on the same NZ x NX cell grid with operator-dependent ("black box") interpolation and
// Galerkin coarse operators, so that 1e16+ conductance contrast (LOW-MASK contact cells,
// APERTURE_MAP_FLOW.F:478-479) does not stall the cycle.  tools/proto_boxmg.py is the numpy
// prototype these kernels restate on the device.
//
//   coarse points      = fine points with even (z,x);  nzc = (nz+1)/2, nxc = (nx+1)/2
//   h-edge  (2Z,2X+1)  : w = (column sums of the stencil) / (S + all couplings but g_N, g_S)
//   v-edge  (2Z+1,2X)  : w = (row sums of the stencil)    / (S + all couplings but g_E, g_W)
//   centre  (2Z+1,2X+1): w = (g_corner + g_vert * w_hedge + g_horz * w_vedge) / diag
//   A_c = P^T A P      : 9-point, stored symmetrically in difference form (S, E, N, NE, NW; C = S + sum g)
// Nothing here subtracts conductances: denominators are sums of non-negative terms, the weight
// deficit 1 - sum(w) is propagated analytically (S/den), and the coarse absorption is
// S_c = P^T (A s) with s = 1 - deficit, evaluated in difference form.  With a subtractive formulation
// the leak conductance of a closed pocket (1e-14 of its internal couplings) drowns in rounding and
// the coarse operators become indefinite (tools/proto_gth.py).
#include "device_utils.cuh"

#include <cstdlib>

namespace lcl {

// Galerkin coarsening next to contact regions produces a few rows that are no M-matrix rows (negative
// couplings = positive off-diagonals).  Taken at face value they give interpolation weights outside
// [0,1], row sums of -8, or a non-positive denominator (then the point got no interpolation at all: a
// hole in P, one slow mode each).  As in classical AMG only the positive couplings interpolate; the
// negative ones are left out of numerator and denominator alike, so weights stay in [0,1] and sum to
// 1 - deficit.  On M-matrix rows (the whole fine level) nothing changes.
__device__ __forceinline__ double pos(double g) { return g > 0.0 ? g : 0.0; }
__device__ __forceinline__ double neg(double g) { return g < 0.0 ? g : 0.0; }

// ------------------------------------------------------------------------------ edge weights
template <class Op>
__global__ void __launch_bounds__(kBlock)
k_weights_edges(Op A, int nzc, int nxc, Weights w, int czlo, int czhi)
{
    const int nz = A.nz(), nx = A.nx();
    const Tiles tiles(czlo, czhi, nxc);
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int Z, X;
        tiles.cell(t, Z, X);
        if (Z >= czhi || X >= nxc) continue;
        const size_t b = (size_t)Z * nxc + X;
        double hw = 0.0, he = 0.0, vs = 0.0, vn = 0.0, th = 0.0, tv = 0.0;
        {   // h-edge point (2Z, 2X+1)
            const int z = 2 * Z, x = 2 * X + 1;
            if (x < nx) {
                const double west = pos(A.g(z, x, 0, -1)) + pos(A.g(z, x, -1, -1)) + pos(A.g(z, x, 1, -1));
                const double east = (x + 1 < nx) ? pos(A.g(z, x, 0, 1)) + pos(A.g(z, x, -1, 1)) + pos(A.g(z, x, 1, 1)) : 0.0;
                const double sp = pos(A.S(z, x));
                const double den = (west + east) + sp;
                if (den > 0.0) { hw = west / den; he = east / den; th = sp / den; }
                else th = 1.0;
            }
        }
        {   // v-edge point (2Z+1, 2X)
            const int z = 2 * Z + 1, x = 2 * X;
            if (z < nz) {
                const double south = pos(A.g(z, x, -1, 0)) + pos(A.g(z, x, -1, -1)) + pos(A.g(z, x, -1, 1));
                const double north = (z + 1 < nz) ? pos(A.g(z, x, 1, 0)) + pos(A.g(z, x, 1, -1)) + pos(A.g(z, x, 1, 1)) : 0.0;
                const double sp = pos(A.S(z, x));
                const double den = (south + north) + sp;
                if (den > 0.0) { vs = south / den; vn = north / den; tv = sp / den; }
                else tv = 1.0;
            }
        }
        w.hw[b] = hw; w.he[b] = he; w.vs[b] = vs; w.vn[b] = vn; w.th[b] = th; w.tv[b] = tv;
    }
}

// ---------------------------------------------------------------------------- centre weights
template <class Op>
__global__ void __launch_bounds__(kBlock)
k_weights_centres(Op A, int nzc, int nxc, Weights w, int czlo, int czhi)
{
    const int nz = A.nz(), nx = A.nx();
    const Tiles tiles(czlo, czhi, nxc);
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int Z, X;
        tiles.cell(t, Z, X);
        if (Z >= czhi || X >= nxc) continue;
        const size_t b = (size_t)Z * nxc + X;
        double c[4] = {0.0, 0.0, 0.0, 0.0};   // sw, se, nw, ne
        double tc = 0.0;
        const int z = 2 * Z + 1, x = 2 * X + 1;
        if (z < nz && x < nx) {
            // diagonal with the negative couplings taken out again: S + (sum of the positive couplings)
            double dg = A.diag(z, x);
#pragma unroll
            for (int dz = -1; dz <= 1; dz++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++)
                    if (dz != 0 || dx != 0) dg -= neg(A.g(z, x, dz, dx));
            dg -= neg(A.S(z, x));
            if (dg > 0.0) {
                // deficit: own absorption + what the four edge neighbours already lack
                double def = pos(A.S(z, x));
                def += pos(A.g(z, x, -1, 0)) * w.th[(size_t)Z * nxc + X];
                if (z + 1 < nz) def += pos(A.g(z, x, 1, 0)) * w.th[(size_t)(Z + 1) * nxc + X];
                def += pos(A.g(z, x, 0, -1)) * w.tv[(size_t)Z * nxc + X];
                if (x + 1 < nx) def += pos(A.g(z, x, 0, 1)) * w.tv[(size_t)Z * nxc + X + 1];
                tc = def / dg;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int cz = (k & 2) ? 1 : -1, cx = (k & 1) ? 1 : -1;
                    if (z + cz >= nz || x + cx >= nx) {
                        // the corner coarse point does not exist: what would have gone there is lost
                        // to the deficit only through the edge neighbours' own (zero) weights
                        continue;
                    }
                    double acc = pos(A.g(z, x, cz, cx));
                    {   // via the h-edge point (z+cz, x): block row (cz<0 ? Z : Z+1), block col X
                        const size_t hb = (size_t)(cz < 0 ? Z : Z + 1) * nxc + X;
                        acc += pos(A.g(z, x, cz, 0)) * (cx < 0 ? w.hw[hb] : w.he[hb]);
                    }
                    {   // via the v-edge point (z, x+cx): block row Z, block col (cx<0 ? X : X+1)
                        const size_t vb = (size_t)Z * nxc + (cx < 0 ? X : X + 1);
                        acc += pos(A.g(z, x, 0, cx)) * (cz < 0 ? w.vs[vb] : w.vn[vb]);
                    }
                    c[k] = acc / dg;
                }
            } else {
                tc = 1.0;
            }
        }
        w.csw[b] = c[0]; w.cse[b] = c[1]; w.cnw[b] = c[2]; w.cne[b] = c[3]; w.tc[b] = tc;
    }
}

// ------------------------------------------------------------------------------ Galerkin RAP
// thread per coarse point I: couplings -A_c(I,J) = -sum_p sum_q P(p,I) A(p,q) P(q,J) for J in
// {E,N,NE,NW}, and the absorption S_c(I) = sum_p P(p,I) (A s)_p with s = 1 - deficit in difference
// form, (A s)_p = S_p s_p + sum_q g_pq (t_q - t_p).  The diagonal follows in k_coarse_diag.
template <class Op>
__global__ void __launch_bounds__(kBlock)
k_galerkin(Op A, int nzc, int nxc, Weights w, double *__restrict__ Sc, double *__restrict__ Ec,
           double *__restrict__ Nc, double *__restrict__ NEc, double *__restrict__ NWc, int czlo, int czhi)
{
    const int nz = A.nz(), nx = A.nx();
    const Tiles tiles(czlo, czhi, nxc);
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int Z, X;
        tiles.cell(t, Z, X);
        if (Z >= czhi || X >= nxc) continue;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        double sabs = 0.0;
        constexpr int jz[4] = {0, 1, 1, 1}, jx[4] = {1, 0, 1, -1};   // E, N, NE, NW of (Z, X)
        // every loop runs over offsets relative to (2Z, 2X) and is fully unrolled, so which weights exist
        // for which (p, q, J) is decided at compile time
#pragma unroll
        for (int a = -1; a <= 1; a++) {
            const int pz = 2 * Z + a;
            if (pz < 0 || pz >= nz) continue;
#pragma unroll
            for (int bb = -1; bb <= 1; bb++) {
                const int px = 2 * X + bb;
                if (px < 0 || px >= nx) continue;
                const double wI = pweight_rel(w, nzc, nxc, Z, X, a, bb, 0, 0);
                if (wI == 0.0) continue;
                const double tp = pdeficit_rel(w, nxc, Z, X, a, bb);
                double as = A.S(pz, px) * (1.0 - tp);
#pragma unroll
                for (int dz = -1; dz <= 1; dz++) {
                    const int qz = pz + dz;
                    if (qz < 0 || qz >= nz) continue;
#pragma unroll
                    for (int dx = -1; dx <= 1; dx++) {
                        const int qx = px + dx;
                        if (qx < 0 || qx >= nx) continue;
                        const bool self = (dz == 0 && dx == 0);
                        const double gpq = self ? 0.0 : A.g(pz, px, dz, dx);
                        const double av = self ? A.diag(pz, px) : -gpq;
                        if (av == 0.0) continue;
                        if (!self) as += gpq * (pdeficit_rel(w, nxc, Z, X, a + dz, bb + dx) - tp);
                        const double wa = wI * av;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            acc[j] += wa * pweight_rel(w, nzc, nxc, Z, X, a + dz, bb + dx, jz[j], jx[j]);
                    }
                }
                sabs += wI * as;
            }
        }
        const size_t i = (size_t)Z * nxc + X;
        Sc[i] = sabs > 0.0 ? sabs : 0.0;
        Ec[i] = (X + 1 < nxc) ? -acc[0] : 0.0;
        Nc[i] = (Z + 1 < nzc) ? -acc[1] : 0.0;
        NEc[i] = (Z + 1 < nzc && X + 1 < nxc) ? -acc[2] : 0.0;
        NWc[i] = (Z + 1 < nzc && X > 0) ? -acc[3] : 0.0;
    }
}

// C = S + sum of the eight couplings (needs every point's E,N,NE,NW, hence a second pass)
__global__ void __launch_bounds__(kBlock)
k_coarse_diag(Coarse9 v, double *__restrict__ C)
{
    const C9 A{v};
    const Tiles tiles(v.zlo, v.zhi, v.nx);
    for (long long t = blockIdx.x; t < tiles.ntiles; t += gridDim.x) {
        int z, x;
        tiles.cell(t, z, x);
        if (z >= v.zhi || x >= v.nx) continue;
        double s = A.S(z, x);
#pragma unroll
        for (int dz = -1; dz <= 1; dz++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++)
                if (dz || dx) s += A.g(z, x, dz, dx);
        C[(size_t)z * v.nx + x] = s;
    }
}

// ------------------------------------------------------------------------ coarsest: dense inverse
// One CTA.  Builds S A S (S = diag^-1/2) densely in global memory and inverts it by Gauss-Jordan
// without pivoting (SPD).  inv is [n][n]; the cycle applies x = S inv S b.
__global__ void k_coarse_inverse(Coarse9 v, double *__restrict__ inv, double *__restrict__ scale, int n)
{
    const C9 A{v};
    const int nx = v.nx, nz = v.nz;
    LCL_DYNAMIC_SMEM(double, sh);   // aug[n][2n]
    double *aug = sh;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) scale[i] = rsqrt(A.diag(i / nx, i % nx));
    __syncthreads();
    for (int e = tid; e < n * 2 * n; e += nt) aug[e] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const int z = i / nx, x = i % nx;
        aug[i * 2 * n + i] = 1.0;   // scaled diagonal
        for (int dz = -1; dz <= 1; dz++)
            for (int dx = -1; dx <= 1; dx++) {
                if (dz == 0 && dx == 0) continue;
                const int qz = z + dz, qx = x + dx;
                if (qz < 0 || qz >= nz || qx < 0 || qx >= nx) continue;
                const int j = qz * nx + qx;
                aug[i * 2 * n + j] = -A.g(z, x, dz, dx) * scale[i] * scale[j];
            }
        aug[i * 2 * n + n + i] = 1.0;
    }
    __syncthreads();
    for (int k = 0; k < n; k++) {
        const double piv = aug[k * 2 * n + k];
        __syncthreads();
        for (int j = tid; j < 2 * n; j += nt) aug[k * 2 * n + j] /= piv;
        __syncthreads();
        for (int e = tid; e < n * 2 * n; e += nt) {
            const int i = e / (2 * n), j = e % (2 * n);
            if (i == k) continue;
            const double f = aug[i * 2 * n + k];
            if (f != 0.0 && j != k) aug[e] -= f * aug[k * 2 * n + j];
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt)
            if (i != k) aug[i * 2 * n + k] = 0.0;
        __syncthreads();
    }
    for (int e = tid; e < n * n; e += nt) inv[e] = aug[(e / n) * 2 * n + n + (e % n)];
}

// ------------------------------------------------------------------------------------ host side
constexpr int kCoarsestCells = 48;   // dense inverse in shared memory: 48*96*8 = 36 KB

// Level sizes, and for a striped solve (world > 1) which levels are striped and where the strips are cut.
// Levels 0 .. gather_level-1 are striped, gather_level and everything below it is held whole by every rank
// and cycled redundantly: a level of a million cells is a handful of microseconds per kernel, launch latency
// that no split shortens, while every striped level costs two neighbour synchronisations per visit.  Strip
// boundaries are multiples of 16 rows on every striped 9-point level (the hybrid smoother's tiles must not
// straddle a cut: Jacobi between tiles, Gauss-Seidel inside, exactly the single-GPU iteration) and of two
// rows on the fine level (red/black parity, 2x2 coarse blocks).
void mg_plan(Solver &s)
{
    s.levels.clear();
    Level l0;
    l0.nz = s.nz; l0.nx = s.nx; l0.n = s.n;
    s.levels.push_back(l0);
    while ((long long)s.levels.back().nz * s.levels.back().nx > kCoarsestCells) {
        const Level &f = s.levels.back();
        Level c;
        c.nz = (f.nz + 1) / 2; c.nx = (f.nx + 1) / 2; c.n = (size_t)c.nz * c.nx;
        s.levels.push_back(c);
    }
    const bool single = s.levels.size() == 1;   // tiny map: one dense "level" equal to the fine grid
    if (single) {
        Level c;
        c.nz = s.nz; c.nx = s.nx; c.n = s.n;
        s.levels.push_back(c);
    }
    const int nlev = (int)s.levels.size();
    // the levels one block cycles in one launch (k_tail_cycle): the first coarse level with at most
    // kTailCells cells and everything below it, if that is more than the dense level alone
    s.tail_top = -1;
    s.tail_smem = 0;
    if (!single && !getenv("APM_LCL_NO_TAIL")) {
        int top = 1;
        while (top < nlev && s.levels[top].n > (size_t)kTailCells) top++;
        size_t framed = 0, plain = 0;
        for (int l = top; l < nlev; l++) {
            framed += (size_t)(s.levels[l].nz + 2) * (s.levels[l].nx + 2);
            plain += s.levels[l].n;
        }
        const size_t smem = (2 * framed + plain) * sizeof(double);
        if (top + 1 < nlev && nlev - top <= kTailMaxLevels && (long long)s.levels.back().n <= 64 && smem <= kTailMaxSmem) {
            s.tail_top = top;
            s.tail_smem = smem;
        }
    }
    // ---- strips
    int lg = 0;
    if (s.world > 1 && !single) {
        size_t min_cells = (size_t)1 << 20;
        if (const char *e = getenv("APM_LCL_STRIPE_MIN_CELLS")) min_cells = (size_t)atoll(e);
        const int cap = s.tail_top > 0 ? s.tail_top : nlev - 1;   // the tail and the dense level are never striped
        while (lg < cap && s.levels[lg].n > min_cells) lg++;
    }
    int cut0[kMaxRanks + 1] = {0};
    for (; lg > 0; lg--) {
        const int align = lg >= 2 ? 16 << (lg - 1) : 2;
        bool ok = true;
        for (int k = 0; k <= s.world; k++) {
            long long c = ((long long)k * s.nz / s.world + align / 2) / align * align;
            cut0[k] = k == s.world ? s.nz : (int)c;
        }
        for (int k = 0; k < s.world; k++) ok &= cut0[k + 1] - cut0[k] >= (align > 32 ? align : 32);
        if (ok) break;
    }
    s.gather_level = lg;
    for (int l = 0; l < nlev; l++) {
        Level &L = s.levels[l];
        L.striped = l < lg;
        for (int k = 0; k <= s.world; k++) {
            L.cut[k] = L.striped ? (k == s.world ? L.nz : cut0[k] >> l) : (k == 0 ? 0 : L.nz);
            const int cnz = l + 1 < nlev ? s.levels[l + 1].nz : 0;
            L.ccut[k] = L.striped ? (k == s.world ? cnz : cut0[k] >> (l + 1)) : (k == 0 ? 0 : cnz);
        }
        if (L.striped) {
            L.zlo = L.cut[s.rank]; L.zhi = L.cut[s.rank + 1];
            L.czlo = L.ccut[s.rank]; L.czhi = L.ccut[s.rank + 1];
        } else {
            L.zlo = 0; L.zhi = L.nz;
            L.czlo = 0; L.czhi = l + 1 < nlev ? s.levels[l + 1].nz : 0;
            // a whole level is "cut" so that rank 0 owns everything: the exchange helpers are never called for it
            for (int k = 1; k <= s.world; k++) { L.cut[k] = L.nz; L.ccut[k] = L.czhi; }
        }
    }
    s.zlo = s.levels[0].zlo;
    s.zhi = s.levels[0].zhi;
}

void mg_allocate(Solver &s, Arena &a)
{
    const int nlev = (int)s.levels.size();
    const bool single = nlev == 2 && s.levels[1].n == s.levels[0].n && s.levels[0].n <= (size_t)kCoarsestCells;
    for (int l = 0; l < nlev; l++) {
        Level &L = s.levels[l];
        if (l > 0) {
            double **cp[9] = {&L.C, &L.S, &L.E, &L.N, &L.NE, &L.NW, &L.x, &L.b, &L.r};
            for (auto p : cp) *p = a.take_rows(L.nx, L.nz, L.striped, L.cut, s.rank, s.world);
        }
        if (l + 1 < nlev && !single) {
            const Level &c = s.levels[l + 1];
            double **wp[11] = {&L.w.hw, &L.w.he, &L.w.vs, &L.w.vn, &L.w.csw, &L.w.cse, &L.w.cnw, &L.w.cne,
                               &L.w.th, &L.w.tv, &L.w.tc};
            for (auto p : wp) *p = a.take_rows(c.nx, c.nz, L.striped, L.ccut, s.rank, s.world);
        } else {
            L.w = Weights{};   // the coarsest level; on the fine level of a tiny map it marks "identity transfer"
        }
    }
    s.coarse_n = (int)s.levels.back().n;
    s.coarse_inv = a.take((size_t)s.coarse_n * s.coarse_n);
    s.coarse_scale = a.take(s.coarse_n);
    s.tail = TailArgs{};
    if (s.tail_top > 0) {
        s.tail.nlev = nlev - s.tail_top;
        for (int l = s.tail_top; l < nlev; l++) {
            const Level &L = s.levels[l];
            s.tail.lv[l - s.tail_top] = TailLevel{L.view9(), L.w, L.x, L.b, L.r};
        }
        s.tail.inv = s.coarse_inv;
        s.tail.scale = s.coarse_scale;
    }
}

// copy of the fine 5-point operator as a 9-point one (only for maps with <= kCoarsestCells cells)
__global__ void k_fine_as_coarse(Fine5 v, double *C, double *S, double *E, double *N, double *NE, double *NW)
{
    const F5 A{v};
    const int n = v.nz * v.nx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int z = i / v.nx, x = i % v.nx;
        C[i] = A.diag(z, x);
        S[i] = A.S(z, x);
        E[i] = A.g(z, x, 0, 1);
        N[i] = A.g(z, x, 1, 0);
        NE[i] = 0.0;
        NW[i] = 0.0;
    }
}

// One coarsening step.  On a striped level every rank computes the weights and the Galerkin rows of its own
// coarse rows [czlo, czhi); what the next kernel reads beyond them arrives by halo exchange (weights: two
// coarse rows either side; operators: the full ghost depth, they are read by every kernel of the cycle).
// The first level held whole by every rank gets its operator rows from all ranks (row_broadcast).
template <class Op>
static void build_next(Solver &s, const Op &A, Level &f, Level &c)
{
    const int grid = grid_for(s, f.czhi - f.czlo, c.nx, 8);
    LAUNCH((k_weights_edges<Op>), grid, kBlock, 0, s.stream, A, c.nz, c.nx, f.w, f.czlo, f.czhi);
    if (f.striped) {
        HaloSet hs;
        for (double *p : {f.w.hw, f.w.he, f.w.vs, f.w.vn, f.w.th, f.w.tv}) hs.add(p);
        halo_exchange(s, hs, c.nx, f.ccut, 2, 2, nullptr);
    }
    LAUNCH((k_weights_centres<Op>), grid, kBlock, 0, s.stream, A, c.nz, c.nx, f.w, f.czlo, f.czhi);
    if (f.striped) {
        HaloSet hs;
        for (double *p : {f.w.csw, f.w.cse, f.w.cnw, f.w.cne, f.w.tc}) hs.add(p);
        halo_exchange(s, hs, c.nx, f.ccut, 2, 2, nullptr);
    }
    LAUNCH((k_galerkin<Op>), grid, kBlock, 0, s.stream, A, c.nz, c.nx, f.w, c.S, c.E, c.N, c.NE, c.NW, f.czlo, f.czhi);
    if (f.striped) {
        HaloSet hs;
        for (double *p : {c.S, c.E, c.N, c.NE, c.NW}) hs.add(p);
        if (c.striped) halo_exchange(s, hs, c.nx, c.cut, kHalo, kHalo, nullptr);
        else row_broadcast(s, hs, c.nx, f.ccut, nullptr);
    }
    LAUNCH(k_coarse_diag, grid_for(s, c.zhi - c.zlo, c.nx, 8), kBlock, 0, s.stream, c.view9(), c.C);
    if (c.striped) {
        HaloSet hs;
        hs.add(c.C);
        halo_exchange(s, hs, c.nx, c.cut, kHalo, kHalo, nullptr);
    }
}

void mg_setup(Solver &s)
{
    if (s.levels[0].w.hw == nullptr) {
        Level &c = s.levels[1];
        LAUNCH(k_fine_as_coarse, 1, 64, 0, s.stream, s.view5(), c.C, c.S, c.E, c.N, c.NE, c.NW);
    } else {
        build_next(s, F5{s.view5()}, s.levels[0], s.levels[1]);
        for (size_t l = 1; l + 1 < s.levels.size(); l++)
            build_next(s, C9{s.levels[l].view9()}, s.levels[l], s.levels[l + 1]);
    }
    const Level &last = s.levels.back();
    const int n = (int)last.n;
    const size_t smem = (size_t)n * 2 * n * sizeof(double);
    LAUNCH(k_coarse_inverse, 1, 256, smem, s.stream, last.view9(), s.coarse_inv, s.coarse_scale, n);
}

// forces the kernels of this file into the device's context now (see preload_kernels, api.cu)
void preload_setup()
{
    preload_one((k_weights_edges<F5>));
    preload_one((k_weights_edges<C9>));
    preload_one((k_weights_centres<F5>));
    preload_one((k_weights_centres<C9>));
    preload_one((k_galerkin<F5>));
    preload_one((k_galerkin<C9>));
    preload_one(k_coarse_diag);
    preload_one(k_coarse_inverse);
    preload_one(k_fine_as_coarse);
}

}   // namespace lcl
