// Fused contact dynamics for LARGE worlds: many bodies (6 nb - neq free velocity components up to ~120) and / or many
// contacts (hundreds), ONE CTA PER WORLD.  Same mathematics, same iterates as the one-warp kernel of dsdf_dynsolve.cu
// (PdipmEngine.solve_dynamics, lcp_physics/physics/engines.py:31-83; PDIPM lcp_physics/lcp/solvers/batch.py:70-237;
// implicit backward lcp_physics/lcp/lcp.py:156-213): the multipliers are eliminated through the per-contact block
// structure of F + D^-1 and only the free block M~_FF = Q_FF + sum_c G_c' B_c^-1 G_c is LU-factored (pivoted, in shared
// memory).  What changes is where things live and who works on them:
//   * shared memory holds only what scales with the number of BODIES: the free block (nF x nF), the per-body mass
//     blocks, the nq-sized vectors, the per-body contact lists;
//   * everything that scales with the number of CONTACTS (Jacobian rows, the nine per-row vectors of the interior
//     point) lives in a caller-provided global workspace (L2 resident: ~1.5 KB per contact), so there is no cap on the
//     contact count (the one-warp kernel stops at 64);
//   * per-body contact lists (ascending contact index, so every sum keeps a fixed order: results are reproducible)
//     replace the "loop over all contacts" of the small kernel in G'w and in the assembly of M~_FF: cost
//     O(nF^2 * contacts per body) instead of O(nF^2 * contacts).
// This is the solver SURVEY.md hard part 4 asks for (KKT beyond shared memory / body-pair block sparsity).
#include "dsdf_math.cuh"
#include "dsdf_dense.cuh"
#include "dsdf_steploop.cuh"
#include "dsdf_dyn_common.cuh"
#include "../../include/dsdf_b200.h"

namespace dsdf {

enum { BV_S = 0, BV_Z, BV_RZ, BV_T, BV_DSA, BV_DZA, BV_DS, BV_DZ, BV_D, BV_COUNT };                 // per-row vectors
enum { BS_XY = 0, BS_RXY, BS_DXYA, BS_DXY, BS_BXY, BS_P, BS_RHS, BS_TF, BS_XF, BS_COUNT };           // nq-sized vectors

struct BigLayout {
    int C, per, R, nz, neq, nq, nF, ldF, nb, half, gs;
    // shared memory (doubles), then ints
    size_t oK, oQ, oS, oRed, oI, smem_bytes;
    // global workspace per world (doubles)
    size_t gG, gYG, gA, gDen, gAP, gV, gMu, gE, gH, gCb, ws_doubles;
};

__host__ __device__ inline BigLayout big_layout(int nb, int neq, int C, int fd) {
    BigLayout L;
    L.C = C; L.per = 2 + fd; L.R = C * L.per; L.nz = 6 * nb; L.neq = neq; L.nq = L.nz + neq; L.nb = nb;
    L.nF = L.nz - neq; L.ldF = L.nF | 1; L.half = fd / 2; L.gs = (1 + L.half) * 12;
    size_t o = 0;
    L.oK = o;   o += (size_t)L.nF * L.ldF;
    L.oQ = o;   o += (size_t)nb * 36;
    L.oS = o;   o += (size_t)BS_COUNT * L.nq;
    L.oRed = o; o += 40;
    L.oI = o;   // ints: eq[2 neq] perm[nF+1] fidx[nz] fpos[nz] bstart[nb+1] blist[2C] piv[2]
    o += ((size_t)2 * neq + L.nF + 1 + 2 * L.nz + nb + 1 + 2 * C + 2 + 1) / 2 + 1;
    L.smem_bytes = o * sizeof(double);
    size_t g = 0;
    L.gG = g;   g += (size_t)C * L.gs;
    L.gYG = g;  g += (size_t)C * 12;
    L.gA = g;   g += L.R;
    L.gDen = g; g += C;
    L.gAP = g;  g += (size_t)2 * C * L.half;
    L.gV = g;   g += (size_t)BV_COUNT * L.R;
    L.gMu = g;  g += C;
    L.gE = g;   g += C;
    L.gH = g;   g += C;
    L.gCb = g;  g += C;                    // 2 ints per contact
    L.ws_doubles = g;
    return L;
}

struct BigCtx {
    BigLayout L;
    int nc, ni;
    double *G, *YG, *a, *den, *AP, *V, *mu, *e, *h;       // global workspace
    double *K, *Qb, *Sv, *red;                              // shared
    int *cb;                                                // global
    int *eq, *perm, *fidx, *fpos, *bstart, *blist, *piv;    // shared
    __device__ double* vec(int k) const { return V + (size_t)k * L.R; }
    __device__ double* sv(int k) const { return Sv + (size_t)k * L.nq; }
    __device__ int gidx(int c, int k) const { return 6 * cb[2 * c + (k >= 6)] + (k % 6); }
};

__device__ inline BigCtx big_ctx(double* sm, double* ws, const BigLayout& L) {
    BigCtx c;
    c.L = L;
    c.K = sm + L.oK; c.Qb = sm + L.oQ; c.Sv = sm + L.oS; c.red = sm + L.oRed;
    int* ib = reinterpret_cast<int*>(sm + L.oI);
    c.eq = ib; c.perm = c.eq + 2 * L.neq; c.fidx = c.perm + L.nF + 1; c.fpos = c.fidx + L.nz; c.bstart = c.fpos + L.nz;
    c.blist = c.bstart + L.nb + 1; c.piv = c.blist + 2 * L.C;
    c.G = ws + L.gG; c.YG = ws + L.gYG; c.a = ws + L.gA; c.den = ws + L.gDen; c.AP = ws + L.gAP; c.V = ws + L.gV;
    c.mu = ws + L.gMu; c.e = ws + L.gE; c.h = ws + L.gH; c.cb = reinterpret_cast<int*>(ws + L.gCb);
    c.nc = c.ni = 0;
    return c;
}

#define BIG_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)

__device__ void big_load(BigCtx& c, int w, const double* p, const double* v, const double* mass, const double* Ibody,
                         const double* fric, const double* rest, const double* f, double dtw, const int* count,
                         const int* cbody, const double* cgeo, const int* eq_rows, int maxc, int fd) {
    const BigLayout& L = c.L;
    const int nb = L.nb, nz = L.nz;
    c.nc = min(min(count[w], maxc), L.C);
    c.ni = c.nc * L.per;
    BIG_FOR(i, 2 * L.neq) c.eq[i] = eq_rows[i];
    BIG_FOR(i, nz) c.fpos[i] = 0;
    __syncthreads();
    BIG_FOR(m, L.neq) c.fpos[6 * c.eq[2 * m] + c.eq[2 * m + 1]] = -(m + 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int jf = 0;
        for (int i = 0; i < nz; ++i) if (c.fpos[i] == 0) { c.fidx[jf] = i; c.fpos[i] = jf++; }
    }
    BIG_FOR(b, nb) {
        const double* pb = p + ((size_t)w * nb + b) * 7;
        M3<double> Iw = winertia<double>(q4<double>(pb[0], pb[1], pb[2], pb[3]), Ibody + ((size_t)w * nb + b) * 9);
        const double m = mass[(size_t)w * nb + b];
        double* Q = c.Qb + 36 * b;
        for (int e = 0; e < 36; ++e) Q[e] = 0.0;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) Q[6 * i + j] = Iw.m[3 * i + j];
            Q[6 * (3 + i) + 3 + i] = m;
        }
    }
    BIG_FOR(cc, c.nc) {
        const size_t oo = (size_t)w * maxc + cc;
        const int i1 = cbody[2 * oo], i2 = cbody[2 * oo + 1];
        c.cb[2 * cc] = i1; c.cb[2 * cc + 1] = i2;
        const double* g = cgeo + 10 * oo;
        const V3<double> n = v3<double>(g[0], g[1], g[2]), p1 = v3<double>(g[3], g[4], g[5]), p2 = v3<double>(g[6], g[7], g[8]);
        double* Gc = c.G + (size_t)cc * L.gs;
        row12<double>(p1, p2, n, Gc);
        V3<double> dirs[8];
        fdirs<double>(n, fd, dirs);
        for (int r = 0; r < L.half; ++r) row12<double>(p1, p2, dirs[r], Gc + 12 * (1 + r));
        c.mu[cc] = 0.5 * (fric[(size_t)w * nb + i1] + fric[(size_t)w * nb + i2]);
        c.e[cc] = (rest[(size_t)w * nb + i1] + rest[(size_t)w * nb + i2]) / 2;
    }
    __syncthreads();
    // per-body contact lists: entry = 2 * contact + side, ascending contact index (one thread per body)
    BIG_FOR(b, nb) {
        int n = 0;
        for (int cc = 0; cc < c.nc; ++cc) n += (c.cb[2 * cc] == b) + (c.cb[2 * cc + 1] == b);
        c.bstart[b + 1] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        c.bstart[0] = 0;
        for (int b = 0; b < nb; ++b) c.bstart[b + 1] += c.bstart[b];
    }
    __syncthreads();
    BIG_FOR(b, nb) {
        int o = c.bstart[b];
        for (int cc = 0; cc < c.nc; ++cc) {
            if (c.cb[2 * cc] == b) c.blist[o++] = 2 * cc;
            if (c.cb[2 * cc + 1] == b) c.blist[o++] = 2 * cc + 1;
        }
    }
    double* pv = c.sv(BS_P);
    BIG_FOR(i, nz) {
        const int b = i / 6, k = i % 6;
        const double* Q = c.Qb + 36 * b;
        double acc = 0.0;
        for (int j = 0; j < 6; ++j) acc += Q[6 * k + j] * v[(size_t)w * nz + 6 * b + j];
        pv[i] = acc + dtw * f[(size_t)w * nz + i];
    }
    BIG_FOR(cc, c.nc) {
        const double* Gc = c.G + (size_t)cc * L.gs;
        double jv = 0.0;
        for (int k = 0; k < 12; ++k) jv += Gc[k] * v[(size_t)w * nz + c.gidx(cc, k)];
        c.h[cc] = jv * c.e[cc];
    }
    __syncthreads();
}

__device__ __forceinline__ double big_Fz_row(const BigCtx& c, const double* z, int r, int fd) {
    const int per = c.L.per, cc = r / per, j = r % per;
    if (j == 0) return 0.0;
    if (j <= fd) return z[cc * per + per - 1];
    double acc = c.mu[cc] * z[cc * per];
    for (int q = 1; q <= fd; ++q) acc -= z[cc * per + q];
    return acc;
}
__device__ __forceinline__ void big_Gx_all(const BigCtx& c, const double* x, double* out) {
    const int per = c.L.per, half = c.L.half, rows = 1 + half;
    BIG_FOR(e, c.nc * rows) {
        const int cc = e / rows, j = e % rows;
        const double* Gc = c.G + (size_t)cc * c.L.gs + 12 * j;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 12; ++k) acc += Gc[k] * x[c.gidx(cc, k)];
        out[cc * per + j] = acc;
        if (j >= 1) out[cc * per + j + half] = -acc;
        else out[cc * per + per - 1] = 0.0;
    }
    __syncthreads();
}
// (G' w)[I] through the contact list of I's body (ascending contact index: same summation order as the small kernel)
__device__ __forceinline__ double big_Gtw_row(const BigCtx& c, const double* w, int I) {
    const int per = c.L.per, half = c.L.half, b = I / 6, k6 = I % 6;
    double acc = 0.0;
    for (int t = c.bstart[b]; t < c.bstart[b + 1]; ++t) {
        const int cc = c.blist[t] >> 1, k = (c.blist[t] & 1) * 6 + k6;
        const double* Gc = c.G + (size_t)cc * c.L.gs + k;
        const double* wc = w + cc * per;
        double a2 = Gc[0] * wc[0];
        for (int q = 1; q <= half; ++q) a2 += Gc[12 * q] * (wc[q] - wc[q + half]);
        acc += a2;
    }
    return acc;
}
__device__ __forceinline__ void big_block_solve(const BigCtx& c, double* t, int fd) {
    const int per = c.L.per;
    BIG_FOR(cc, c.nc) {
        double* tc = t + cc * per;
        const double* ac = c.a + cc * per;
        const double un = tc[0] * ac[0];
        double sum = 0.0;
        for (int q = 1; q <= fd; ++q) sum += tc[q] * ac[q];
        const double ug = (tc[per - 1] - c.mu[cc] * un + sum) * c.den[cc];
        tc[0] = un;
        for (int q = 1; q <= fd; ++q) tc[q] = (tc[q] - ug) * ac[q];
        tc[per - 1] = ug;
    }
    __syncthreads();
}

__device__ int big_factor(BigCtx& c, const double* d, int fd) {
    const BigLayout& L = c.L;
    const int per = L.per, nF = L.nF, ld = L.ldF, half = L.half;
    BIG_FOR(r, c.ni) c.a[r] = 1.0 / (1.0 / d[r]);
    __syncthreads();
    BIG_FOR(cc, c.nc) {
        const double* ac = c.a + cc * per;
        double dn = 1.0 / d[cc * per + per - 1];
        for (int q = 1; q <= fd; ++q) dn += ac[q];
        c.den[cc] = 1.0 / dn;
    }
    BIG_FOR(e, c.nc * half) {
        const int cc = e / half, q = 1 + e % half;
        const double* ac = c.a + cc * per;
        c.AP[(size_t)2 * e] = ac[q] + ac[q + half];
        c.AP[(size_t)2 * e + 1] = ac[q] - ac[q + half];
    }
    __syncthreads();
    BIG_FOR(e, c.nc * 12) {
        const int cc = e / 12, k = e % 12;
        const double* Gc = c.G + (size_t)cc * L.gs + k;
        const double* ap = c.AP + (size_t)2 * cc * half;
        double acc = -c.mu[cc] * Gc[0] * c.a[cc * per];
        for (int q = 1; q <= half; ++q) acc += Gc[12 * q] * ap[2 * (q - 1) + 1];
        c.YG[e] = acc * c.den[cc];
    }
    __syncthreads();
    // M~_FF: entry (If, jf) gathers the contacts that touch BOTH bodies, walking the list of I's body
    BIG_FOR(e, nF * nF) {
        const int If = e / nF, jf = e % nF, I = c.fidx[If], J = c.fidx[jf], bI = I / 6, bJ = J / 6;
        double acc = bI == bJ ? c.Qb[36 * bI + 6 * (I % 6) + (J % 6)] : 0.0;
        for (int t = c.bstart[bI]; t < c.bstart[bI + 1]; ++t) {
            const int cc = c.blist[t] >> 1, sI = c.blist[t] & 1;
            int kj;
            if (bJ == bI) kj = sI * 6 + J % 6;
            else if (c.cb[2 * cc + (1 - sI)] == bJ) kj = (1 - sI) * 6 + J % 6;
            else continue;
            const int ki = sI * 6 + I % 6;
            const double* Gc = c.G + (size_t)cc * L.gs;
            const double* ap = c.AP + (size_t)2 * cc * half;
            double a2 = Gc[ki] * (Gc[kj] * c.a[cc * per]);
            const double yg = c.YG[cc * 12 + kj];
            for (int q = 1; q <= half; ++q)
                a2 += Gc[12 * q + ki] * (Gc[12 * q + kj] * ap[2 * (q - 1)] - yg * ap[2 * (q - 1) + 1]);
            acc += a2;
        }
        c.K[If * ld + jf] = acc;
    }
    if (threadIdx.x == 0) c.piv[1] = 0;
    __syncthreads();
    block_lu(c.K, ld, nF, c.perm, &c.piv[0], &c.piv[1]);
    __syncthreads();
    return c.piv[1];
}

__device__ inline void big_kkt_solve(BigCtx& c, const double* rhs, double* dxy) {
    const BigLayout& L = c.L;
    const int nz = L.nz, nF = L.nF, ld = L.ldF;
    double *tF = c.sv(BS_TF), *xF = c.sv(BS_XF);
    BIG_FOR(jf, nF) tF[jf] = rhs[c.fidx[jf]];
    __syncthreads();
    block_lu_solve(c.K, ld, nF, c.perm, tF, xF);
    BIG_FOR(I, nz) { const int fp = c.fpos[I]; dxy[I] = fp >= 0 ? xF[fp] : rhs[nz + (-fp - 1)]; }
    __syncthreads();
}

// batch.py:380-410 semantics, see dyn_solve in dsdf_dynsolve.cu
__device__ void big_solve(BigCtx& c, const double* d, int fd, const double* rxy, const double* rs, const double* rz,
                          double* dxy, double* ds, double* dz) {
    const BigLayout& L = c.L;
    const int nz = L.nz, nq = L.nq;
    double* t = c.vec(BV_T);
    double* rhs = c.sv(BS_RHS);
    BIG_FOR(r, c.ni) { const double tv = (rz ? rz[r] : 0.0) - (rs ? rs[r] / d[r] : 0.0); t[r] = tv; dz[r] = tv; }
    __syncthreads();
    big_block_solve(c, t, fd);
    BIG_FOR(I, nq) {
        double acc = rxy ? -rxy[I] : 0.0;
        if (I < nz) acc -= big_Gtw_row(c, t, I);
        rhs[I] = acc;
    }
    __syncthreads();
    big_kkt_solve(c, rhs, dxy);
    double* gx = t;
    big_Gx_all(c, dxy, gx);
    BIG_FOR(r, c.ni) dz[r] = gx[r] + dz[r];
    __syncthreads();
    big_block_solve(c, dz, fd);
    BIG_FOR(m, L.neq) {
        const int b = c.eq[2 * m], k = c.eq[2 * m + 1], I = 6 * b + k;
        double acc = (rxy ? -rxy[I] : 0.0) - big_Gtw_row(c, dz, I);
        for (int j = 0; j < 6; ++j) acc -= c.Qb[36 * b + 6 * k + j] * dxy[6 * b + j];
        dxy[nz + m] = acc;
    }
    BIG_FOR(r, c.ni) ds[r] = ((rs ? -rs[r] : 0.0) - dz[r]) / d[r];
    __syncthreads();
}

__device__ double big_ratio_step(const BigCtx& c, const double* v, const double* dv) {
    double mx = -INFINITY;
    BIG_FOR(r, c.ni) mx = fmax(mx, -v[r] / dv[r]);
    mx = block_reduce<RED_MAX>(mx, c.red);
    const double repl = fmax(1.0, mx);
    double mn = INFINITY;
    BIG_FOR(r, c.ni) mn = fmin(mn, dv[r] > 0.0 ? repl : -v[r] / dv[r]);
    return block_reduce<RED_MIN>(mn, c.red);
}

#ifndef DSDF_BIG_THREADS
#define DSDF_BIG_THREADS 256
#endif

__global__ void __launch_bounds__(DSDF_BIG_THREADS)
dyn_forward_big_kernel(const double* __restrict__ p, const double* __restrict__ v, const double* __restrict__ mass,
                       const double* __restrict__ Ibody, const double* __restrict__ fric, const double* __restrict__ rest,
                       const double* __restrict__ f, const double* __restrict__ dt, const unsigned char* __restrict__ active,
                       const int* __restrict__ count, const int* __restrict__ cbody, const double* __restrict__ cgeo,
                       const int* __restrict__ eq_rows, int nb, int neq, int maxc, int C, int fd,
                       double eps, int not_improved_lim, int max_iter,
                       double* __restrict__ xo, double* __restrict__ nvo, double* __restrict__ nuo, double* __restrict__ lamo,
                       double* __restrict__ so, int* __restrict__ status_o, int* __restrict__ iters_o,
                       const int* __restrict__ vmap, int* __restrict__ ctrl, int cmin, int last_class,
                       double* __restrict__ ws_all) {
    extern __shared__ double smd[];
    const int w = blockIdx.x, tid = threadIdx.x;
    if (ctrl && (loop_idle(ctrl) || w >= ctrl[CT_NVIRT])) return;
    const int wsrc = vmap ? vmap[w] : w;
    const BigLayout L = big_layout(nb, neq, C, fd);
    const int nz = L.nz, nq = L.nq, per = L.per;
    if (active && !active[w]) {
        if (nvo) BIG_FOR(i, nz) nvo[(size_t)w * nz + i] = v[(size_t)wsrc * nz + i];
        return;
    }
    if (count[wsrc] <= cmin || (count[wsrc] > C && !last_class)) return;
    if (count[wsrc] > C) {
        BIG_FOR(i, nz) { xo[(size_t)w * nz + i] = NAN; if (nvo) nvo[(size_t)w * nz + i] = NAN; }
        if (tid == 0) {
            status_o[w] = DSDF_LCP_TOO_LARGE; if (iters_o) iters_o[w] = 0;
            if (ctrl) atomicOr(&ctrl[CT_ABORT], DSDF_STEP_DYN_SMEM);
        }
        return;
    }
    BigCtx c = big_ctx(smd, ws_all + (size_t)w * L.ws_doubles, L);
    big_load(c, wsrc, p, v, mass, Ibody, fric, rest, f, dt[w], count, cbody, cgeo, eq_rows, maxc, fd);
    auto hrow = [&](int r) { return (r % per == 0) ? c.h[r / per] : 0.0; };
    const int niCap = maxc * per;
    const int ni = c.ni;
    BIG_FOR(r, niCap) { lamo[(size_t)w * niCap + r] = 0.0; so[(size_t)w * niCap + r] = 0.0; }
    double *s = c.vec(BV_S), *z = c.vec(BV_Z), *rz = c.vec(BV_RZ), *dsa = c.vec(BV_DSA), *dza = c.vec(BV_DZA),
           *ds = c.vec(BV_DS), *dz = c.vec(BV_DZ), *d = c.vec(BV_D);
    double *xy = c.sv(BS_XY), *rxy = c.sv(BS_RXY), *dxya = c.sv(BS_DXYA), *dxy = c.sv(BS_DXY), *bxy = c.sv(BS_BXY),
           *pv = c.sv(BS_P);
    int status = 0, iters = 0;
    bool have_best = false;
    double best_res = INFINITY, mu_gap = 0.0, sz = 0.0;
    int stalled = 0;
    for (int it = -1; it < max_iter; ++it) {
        double res = 0.0;
        if (it < 0) {
            BIG_FOR(r, ni) { d[r] = 1.0; rz[r] = -hrow(r); }
            BIG_FOR(i, nq) rxy[i] = i < nz ? pv[i] : 0.0;
        } else {
            double nrx = 0.0, nry = 0.0, nrz = 0.0;
            sz = 0.0;
            BIG_FOR(I, nq) {
                double acc;
                if (I < nz) {
                    const int b = I / 6, k = I % 6;
                    const int fp = c.fpos[I];
                    const double ay = fp < 0 ? xy[nz + (-fp - 1)] : 0.0;
                    double qx = 0.0;
                    for (int j = 0; j < 6; ++j) qx += xy[6 * b + j] * c.Qb[36 * b + 6 * k + j];
                    acc = ay + big_Gtw_row(c, z, I) + qx + pv[I];
                    nrx += acc * acc;
                } else {
                    const int m = I - nz;
                    acc = xy[6 * c.eq[2 * m] + c.eq[2 * m + 1]];
                    nry += acc * acc;
                }
                rxy[I] = acc;
            }
            big_Gx_all(c, xy, c.vec(BV_T));
            BIG_FOR(r, ni) {
                const double val_ = c.vec(BV_T)[r] + s[r] - hrow(r) - big_Fz_row(c, z, r, fd);
                rz[r] = val_;
                nrz += val_ * val_;
                sz += s[r] * z[r];
            }
            nrx = block_reduce<RED_SUM>(nrx, c.red); nry = block_reduce<RED_SUM>(nry, c.red);
            nrz = block_reduce<RED_SUM>(nrz, c.red); sz = block_reduce<RED_SUM>(sz, c.red);
            mu_gap = fabs(sz / ni);
            res = (L.neq > 0 ? sqrt(nry) : 0.0) + sqrt(nrz) + sqrt(nrx) + ni * mu_gap;
            BIG_FOR(r, ni) d[r] = z[r] / s[r];
        }
        __syncthreads();
        if (big_factor(c, d, fd)) { status |= DSDF_LCP_FACTOR_FAIL; if (it >= 0) break; }
        if (it >= 0) {
            iters = it + 1;
            if (!have_best || res < best_res) {
                have_best = true; best_res = res; stalled = 0;
                BIG_FOR(i, nq) bxy[i] = xy[i];
                BIG_FOR(r, ni) {
                    const int rr = ref_row(r / per, r % per, c.nc, fd);
                    lamo[(size_t)w * niCap + rr] = z[r];
                    so[(size_t)w * niCap + rr] = s[r];
                }
            } else {
                ++stalled;
            }
            if (stalled == not_improved_lim || best_res < eps || mu_gap > 1e32) break;
        }
        bool init_done = false;
        for (int pass = 0; pass < 2; ++pass) {
            const double* a_rxy = pass == 0 ? rxy : nullptr;
            const double* a_rs = pass == 0 ? (it < 0 ? nullptr : z) : rz;
            const double* a_rz = pass == 0 ? rz : nullptr;
            double* o_xy = pass == 0 ? (it < 0 ? xy : dxya) : dxy;
            double* o_s = pass == 0 ? (it < 0 ? s : dsa) : ds;
            double* o_z = pass == 0 ? (it < 0 ? z : dza) : dz;
            big_solve(c, d, fd, a_rxy, a_rs, a_rz, o_xy, o_s, o_z);
            if (it < 0) {
                if (ni > 0) {
                    double ms = INFINITY, mz = INFINITY;
                    BIG_FOR(r, ni) { ms = fmin(ms, s[r]); mz = fmin(mz, z[r]); }
                    ms = block_reduce<RED_MIN>(ms, c.red); mz = block_reduce<RED_MIN>(mz, c.red);
                    BIG_FOR(r, ni) {
                        if (ms < 0.0) s[r] -= ms - 1.0;
                        if (mz < 0.0) z[r] -= mz - 1.0;
                    }
                    __syncthreads();
                }
                init_done = true;
                break;
            }
            if (pass == 0) {
                double alpha = fmin(fmin(big_ratio_step(c, z, dza), big_ratio_step(c, s, dsa)), 1.0);
                double t3 = 0.0;
                BIG_FOR(r, ni) t3 += (s[r] + alpha * dsa[r]) * (z[r] + alpha * dza[r]);
                t3 = block_reduce<RED_SUM>(t3, c.red);
                const double r3 = t3 / sz, sig = r3 * r3 * r3;
                BIG_FOR(r, ni) rz[r] = (-mu_gap * sig + dsa[r] * dza[r]) / s[r];
                __syncthreads();
            }
        }
        if (init_done) {
            if (ni == 0) {
                BIG_FOR(i, nq) bxy[i] = xy[i];
                have_best = true;
                break;
            }
            continue;
        }
        BIG_FOR(i, nq) dxy[i] += dxya[i];
        BIG_FOR(r, ni) { ds[r] += dsa[r]; dz[r] += dza[r]; }
        __syncthreads();
        const double alpha = fmin(0.999 * fmin(big_ratio_step(c, z, dz), big_ratio_step(c, s, ds)), 1.0);
        BIG_FOR(i, nq) xy[i] += alpha * dxy[i];
        BIG_FOR(r, ni) { s[r] += alpha * ds[r]; z[r] += alpha * dz[r]; }
        __syncthreads();
    }
    if (ni > 0 && have_best && best_res > 1.0) status |= DSDF_LCP_INACCURATE;
    __syncthreads();
    BIG_FOR(i, nz) {
        xo[(size_t)w * nz + i] = have_best ? bxy[i] : NAN;
        if (nvo) nvo[(size_t)w * nz + i] = have_best ? -bxy[i] : NAN;
    }
    BIG_FOR(m, L.neq) nuo[(size_t)w * L.neq + m] = have_best ? bxy[nz + m] : NAN;
    if (tid == 0) { status_o[w] = status; if (iters_o) iters_o[w] = iters; }
}

// implicit differentiation at the solution, contracted onto the physical inputs (see dyn_backward_kernel)
__global__ void __launch_bounds__(DSDF_BIG_THREADS)
dyn_backward_big_kernel(const double* __restrict__ p, const double* __restrict__ v, const double* __restrict__ mass,
                        const double* __restrict__ Ibody, const double* __restrict__ fric, const double* __restrict__ rest,
                        const double* __restrict__ f, const double* __restrict__ dt, const unsigned char* __restrict__ active,
                        const int* __restrict__ count, const int* __restrict__ cbody, const double* __restrict__ cgeo,
                        const int* __restrict__ eq_rows, int nb, int neq, int maxc, int C, int fd,
                        int stop_contact_grad, int stop_friction_grad,
                        const double* __restrict__ xs, const double* __restrict__ lams, const double* __restrict__ ss,
                        const double* __restrict__ gnv,
                        double* __restrict__ gp, double* __restrict__ gv, double* __restrict__ gmass, double* __restrict__ gI,
                        double* __restrict__ gfric, double* __restrict__ grest, double* __restrict__ gf,
                        double* __restrict__ gdt, double* __restrict__ ggeo, int cmin, int last_class,
                        double* __restrict__ ws_all) {
    extern __shared__ double smd[];
    const int w = blockIdx.x, tid = threadIdx.x;
    const BigLayout L = big_layout(nb, neq, C, fd);
    const int nz = L.nz, nq = L.nq, per = L.per, niCap = maxc * per;
    const bool masked = active && !active[w];
    if (masked ? cmin >= 0 : (count[w] <= cmin || (count[w] > C && !last_class))) return;
    const bool on = !masked && count[w] <= C;
    if (!on) {
        BIG_FOR(i, nb * 7) gp[(size_t)w * nb * 7 + i] = 0.0;
        BIG_FOR(i, nz) { gv[(size_t)w * nz + i] = masked ? gnv[(size_t)w * nz + i] : 0.0; gf[(size_t)w * nz + i] = 0.0; }
        BIG_FOR(i, nb) { gmass[(size_t)w * nb + i] = 0.0; gfric[(size_t)w * nb + i] = 0.0; grest[(size_t)w * nb + i] = 0.0; }
        BIG_FOR(i, nb * 9) gI[(size_t)w * nb * 9 + i] = 0.0;
        BIG_FOR(i, maxc * 10) ggeo[(size_t)w * maxc * 10 + i] = 0.0;
        if (tid == 0) gdt[w] = 0.0;
        return;
    }
    BigCtx c = big_ctx(smd, ws_all + (size_t)w * L.ws_doubles, L);
    const double dtw = dt[w];
    big_load(c, w, p, v, mass, Ibody, fric, rest, f, dtw, count, cbody, cgeo, eq_rows, maxc, fd);
    const int ni = c.ni, nc = c.nc;
    const double* vw = v + (size_t)w * nz;
    double *lam = c.vec(BV_Z), *d = c.vec(BV_D), *dl = c.vec(BV_DZ), *ds = c.vec(BV_DS);
    double *g = c.sv(BS_RXY), *dxy = c.sv(BS_DXY), *zh = c.sv(BS_XY);
    BIG_FOR(i, nq) { g[i] = i < nz ? -gnv[(size_t)w * nz + i] : 0.0; zh[i] = i < nz ? xs[(size_t)w * nz + i] : 0.0; }
    BIG_FOR(r, ni) {
        const int rr = ref_row(r / per, r % per, nc, fd);
        lam[r] = lams[(size_t)w * niCap + rr];
        d[r] = fmax(lam[r], 1e-8) / fmax(ss[(size_t)w * niCap + rr], 1e-8);
    }
    __syncthreads();
    big_factor(c, d, fd);
    big_solve(c, d, fd, g, nullptr, nullptr, dxy, ds, dl);
    BIG_FOR(I, nz) {
        const int b = I / 6, k = I % 6;
        double acc = 0.0;
        for (int i = 0; i < 6; ++i) acc += c.Qb[36 * b + 6 * i + k] * dxy[6 * b + i];
        for (int t = c.bstart[b]; t < c.bstart[b + 1]; ++t) {
            const int cc = c.blist[t] >> 1, kk = (c.blist[t] & 1) * 6 + k;
            acc += c.G[(size_t)cc * L.gs + kk] * c.e[cc] * (-dl[cc * per]);
        }
        gv[(size_t)w * nz + I] = acc;
        gf[(size_t)w * nz + I] = dtw * dxy[I];
    }
    {
        double acc = 0.0;
        BIG_FOR(I, nz) acc += f[(size_t)w * nz + I] * dxy[I];
        acc = block_reduce<RED_SUM>(acc, c.red);
        if (tid == 0) gdt[w] = acc;
    }
    BIG_FOR(b, nb) {
        double dMb[36];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j)
                dMb[6 * i + j] = 0.5 * (dxy[6 * b + i] * zh[6 * b + j] + zh[6 * b + i] * dxy[6 * b + j]) +
                                 dxy[6 * b + i] * vw[6 * b + j];
        gmass[(size_t)w * nb + b] = dMb[6 * 3 + 3] + dMb[6 * 4 + 4] + dMb[6 * 5 + 5];
        const double* pb = p + ((size_t)w * nb + b) * 7;
        const double* Ib = Ibody + ((size_t)w * nb + b) * 9;
        for (int seed = 0; seed < 13; ++seed) {
            Dual I9[9];
            for (int e = 0; e < 9; ++e) I9[e] = Dual(Ib[e], seed == 4 + e ? 1.0 : 0.0);
            Q4<Dual> q = q4<Dual>(Dual(pb[0], seed == 0), Dual(pb[1], seed == 1), Dual(pb[2], seed == 2), Dual(pb[3], seed == 3));
            M3<Dual> Iw = winertia<Dual>(q, I9);
            double acc = 0.0;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) acc += dMb[6 * i + j] * Iw.m[3 * i + j].d;
            if (seed < 4) gp[((size_t)w * nb + b) * 7 + seed] = acc;
            else gI[((size_t)w * nb + b) * 9 + seed - 4] = acc;
        }
        gp[((size_t)w * nb + b) * 7 + 4] = 0.0; gp[((size_t)w * nb + b) * 7 + 5] = 0.0; gp[((size_t)w * nb + b) * 7 + 6] = 0.0;
        double gfr = 0.0, gre = 0.0;
        for (int t = c.bstart[b]; t < c.bstart[b + 1]; ++t) {
            const int cc = c.blist[t] >> 1;
            gfr += 0.5 * dl[cc * per + per - 1] * lam[cc * per];
            double jv = 0.0;
            for (int k = 0; k < 12; ++k) jv += c.G[(size_t)cc * L.gs + k] * vw[c.gidx(cc, k)];
            gre += 0.5 * (-dl[cc * per]) * jv;
        }
        gfric[(size_t)w * nb + b] = gfr;
        grest[(size_t)w * nb + b] = gre;
    }
    BIG_FOR(t, maxc * 10) {
        const int cc = t / 10, comp = t % 10;
        double acc = 0.0;
        if (cc < nc && comp < 9) {
            const double* gg = cgeo + 10 * ((size_t)w * maxc + cc);
            auto D = [&](int k) { return Dual(gg[k], comp == k ? 1.0 : 0.0); };
            const V3<Dual> n = v3<Dual>(D(0), D(1), D(2)), p1 = v3<Dual>(D(3), D(4), D(5)), p2 = v3<Dual>(D(6), D(7), D(8));
            Dual row[12];
            if (!stop_contact_grad) {
                row12<Dual>(p1, p2, n, row);
                const double dlr = dl[cc * per], lr = lam[cc * per], dh = -dlr;
                for (int k = 0; k < 12; ++k) {
                    const int I = c.gidx(cc, k);
                    acc += (dlr * zh[I] + lr * dxy[I] + dh * c.e[cc] * vw[I]) * row[k].d;
                }
            }
            if (!stop_friction_grad) {
                V3<Dual> dirs[8];
                fdirs<Dual>(n, fd, dirs);
                for (int r = 0; r < fd; ++r) {
                    row12<Dual>(p1, p2, dirs[r], row);
                    const double dlr = dl[cc * per + 1 + r], lr = lam[cc * per + 1 + r];
                    for (int k = 0; k < 12; ++k) {
                        const int I = c.gidx(cc, k);
                        acc += (dlr * zh[I] + lr * dxy[I]) * row[k].d;
                    }
                }
            }
        }
        ggeo[(size_t)w * maxc * 10 + t] = acc;
    }
}

}  // namespace dsdf

using namespace dsdf;

extern "C" {

size_t dsdf_dynamics_big_smem_bytes(int nb, int neq, int ncontacts, int fric_dirs) {
    return big_layout(nb, neq, ncontacts, fric_dirs).smem_bytes;
}
size_t dsdf_dynamics_big_workspace_bytes(int W, int nb, int neq, int ncontacts, int fric_dirs) {
    return (size_t)W * big_layout(nb, neq, ncontacts, fric_dirs).ws_doubles * sizeof(double);
}

static int big_check(int W, int nb, int neq, int maxc, int* C, int fd, size_t* smem, const void* ws) {
    if (W <= 0 || nb <= 0 || neq < 0 || maxc <= 0 || (fd != 8 && fd != 4) || !ws) return -1;
    if (*C <= 0 || *C > maxc) *C = maxc;
    *smem = big_layout(nb, neq, *C, fd).smem_bytes;
    if (*smem > 227 * 1024) return -2;
    return 0;
}
static size_t g_big_fwd = 0, g_big_bwd = 0;

int dsdf_dynamics_big_solve(const double* p, const double* v, const double* mass, const double* Ibody,
                            const double* fric, const double* rest, const double* f, const double* dt,
                            const unsigned char* active, const int32_t* count, const int32_t* cbody, const double* cgeo,
                            const int32_t* eq_rows, int W, int nb, int neq, int maxc, int ncontacts, int fric_dirs,
                            double eps, int not_improved_lim, int max_iter,
                            double* x, double* new_v, double* nu, double* lam, double* s, int32_t* status, int32_t* iters,
                            const int32_t* vmap, int32_t* ctrl, int count_min, int last_class, double* workspace,
                            void* stream) {
    size_t smem;
    int C = ncontacts;
    int rc = big_check(W, nb, neq, maxc, &C, fric_dirs, &smem, workspace);
    if (rc) return rc;
    cudaError_t e = ensure_smem(dyn_forward_big_kernel, smem, &g_big_fwd);
    if (e != cudaSuccess) return (int)e;
    dyn_forward_big_kernel<<<W, DSDF_BIG_THREADS, smem, (cudaStream_t)stream>>>(
        p, v, mass, Ibody, fric, rest, f, dt, active, count, cbody, cgeo, eq_rows, nb, neq, maxc, C, fric_dirs, eps,
        not_improved_lim, max_iter, x, new_v, nu, lam, s, status, iters, vmap, ctrl, count_min, last_class, workspace);
    return (int)cudaGetLastError();
}

int dsdf_dynamics_big_solve_backward(const double* p, const double* v, const double* mass, const double* Ibody,
                                     const double* fric, const double* rest, const double* f, const double* dt,
                                     const unsigned char* active, const int32_t* count, const int32_t* cbody,
                                     const double* cgeo, const int32_t* eq_rows, int W, int nb, int neq, int maxc,
                                     int ncontacts, int fric_dirs, int stop_contact_grad, int stop_friction_grad,
                                     const double* x, const double* lam, const double* s, const double* g_new_v,
                                     double* gp, double* gv, double* gmass, double* gI, double* gfric, double* grest,
                                     double* gf, double* gdt, double* ggeo, int count_min, int last_class,
                                     double* workspace, void* stream) {
    size_t smem;
    int C = ncontacts;
    int rc = big_check(W, nb, neq, maxc, &C, fric_dirs, &smem, workspace);
    if (rc) return rc;
    cudaError_t e = ensure_smem(dyn_backward_big_kernel, smem, &g_big_bwd);
    if (e != cudaSuccess) return (int)e;
    dyn_backward_big_kernel<<<W, DSDF_BIG_THREADS, smem, (cudaStream_t)stream>>>(
        p, v, mass, Ibody, fric, rest, f, dt, active, count, cbody, cgeo, eq_rows, nb, neq, maxc, C, fric_dirs,
        stop_contact_grad, stop_friction_grad, x, lam, s, g_new_v, gp, gv, gmass, gI, gfric, grest, gf, gdt, ggeo,
        count_min, last_class, workspace);
    return (int)cudaGetLastError();
}

}  // extern "C"
