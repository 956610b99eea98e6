// Fused contact dynamics: LCP assembly + primal-dual interior point + (backward) implicit differentiation,
// ONE WARP PER WORLD, exploiting the contact structure of the reference's LCP instead of a dense factorisation.
//
// Replaces PdipmEngine.solve_dynamics (lcp_physics/physics/engines.py:31-83) end to end, i.e. the same mathematics
// as dsdf_dynamics_assemble + dsdf_lcp_forward/backward (which remain the dense, general LCPFunction path):
//   * PDIPM iteration, initial point, best-iterate tracking, step rules: lcp_physics/lcp/solvers/batch.py:70-237
//   * implicit backward: lcp_physics/lcp/lcp.py:156-213
// Linear algebra.  The reference eliminates x first and LU-factors the nineq x nineq matrix
// G Q^-1 G' + F + D^-1 (batch.py:413-520).  In the engine's LCP, F + D^-1 is block diagonal per contact
// (rows: normal, fric_dirs friction rows, one cone row; engines.py:72-78):
//        B_c = [[a_n, 0, 0], [0, diag(a_f), 1], [mu_c, -1', a_g]],   a = 1/d = s/z > 0
// with a closed-form inverse, so we eliminate the multipliers instead and factor the (nz+neq)^2 matrix
//        K = [[Q + sum_c G_c' B_c^-1 G_c, A'], [A, 0]]
// The equality rows of the engine are 0/1 selections of single velocity components with b = 0 (pinned bodies, axis
// locks: sdf_physics/physics3d/constraints.py:32-145), so x_P = 0 on the pinned set P throughout and K reduces once
// more: pivoted LU of the FREE block M~_FF only ((6 nb - neq)^2: 6 x 6 for the box-on-plane scene instead of
// 100 x 100), multipliers dy = r_P - M~_PF dx_F from the pinned rows.  Same Newton systems, same iterates up to
// round-off; verified against the dense kernels and the oracle in tests/test_dynsolve_gpu.py.
#include "dsdf_math.cuh"
#include "dsdf_dense.cuh"
#include "dsdf_steploop.cuh"
#include "dsdf_dyn_common.cuh"
#include "../../include/dsdf_b200.h"

namespace dsdf {

struct DynSmem {
    int C, per, R, nz, neq, nq, nF, ldF, nb, half, gs;   // gs = doubles of G per contact = (1 + fd/2) * 12
    size_t oG, oYG, oA, oDen, oAP, oV, oK, oQ, oS, oI, bytes;
};
enum { DV_S = 0, DV_Z, DV_RZ, DV_T, DV_DSA, DV_DZA, DV_DS, DV_DZ, DV_D, DV_COUNT };   // 9 row vectors per world
enum { DS_XY = 0, DS_RXY, DS_DXYA, DS_DXY, DS_BXY, DS_P, DS_RHS, DS_TF, DS_XF, DS_COUNT };

__host__ __device__ inline DynSmem dyn_layout(int nb, int neq, int C, int fd) {
    DynSmem L;
    L.C = C; L.per = 2 + fd; L.R = C * L.per; L.nz = 6 * nb; L.neq = neq; L.nq = L.nz + neq; L.nb = nb;
    L.nF = L.nz - neq; L.ldF = L.nF | 1; L.half = fd / 2; L.gs = (1 + L.half) * 12;
    size_t o = 0;
    L.oG = o;   o += (size_t)C * L.gs;        // normal row + the first fd/2 friction rows (the rest are their negatives)
    L.oYG = o;  o += (size_t)C * 12;
    L.oA = o;   o += L.R;                 // a = 1/d
    L.oDen = o; o += C;
    L.oAP = o;  o += (size_t)2 * C * L.half;  // per friction pair: 1/a_q + 1/a_{q+half}, 1/a_q - 1/a_{q+half}
    L.oV = o;   o += (size_t)DV_COUNT * L.R;
    L.oK = o;   o += (size_t)L.nz * L.ldF;      // rows [0,nF): M~_FF (LU in place); rows [nF,nz): M~_PF
    L.oQ = o;   o += (size_t)nb * 36;
    L.oS = o;   o += (size_t)DS_COUNT * L.nq;
    L.oI = o;   o += (size_t)(2 * C + 2 * neq + L.nq + 2 * L.nz + 8 + 1) / 2 + 1;
    L.bytes = o * sizeof(double);
    return L;
}

struct DynCtx {
    DynSmem L;
    int nc, ni;                     // active contacts / rows of this world
    double *G, *YG, *a, *den, *AP, *V, *K, *Qb, *Sv;
    int *cb, *eq, *perm, *fidx, *fpos;   // fidx[jf] = free dof; fpos[I] = jf (free) or -(m+1) (pinned by row m)
    __device__ double* vec(int k) const { return V + (size_t)k * L.R; }
    __device__ double* sv(int k) const { return Sv + (size_t)k * L.nq; }
    __device__ int gidx(int c, int k) const { return 6 * cb[2 * c + (k >= 6)] + (k % 6); }
};

// DSDF_DYN_OUTLINE: keep ONE out-of-line copy of the building blocks (the kernel is 17 k SASS instructions = 280 KB when
// everything is inlined, far beyond the instruction cache; ncu: 13-34 % of stall samples "no_instructions")
#ifdef DSDF_DYN_OUTLINE
#define DYN_FN __device__ __noinline__
#else
#define DYN_FN __device__ __forceinline__
#endif
__device__ __forceinline__ double wsum(double v) { return warp_sum(v); }

// ---- warp-level pivoted LU / solve on a small matrix in shared memory ---------------------------------------
DYN_FN int warp_lu(double* A, int ld, int n, int* perm) {
    const int lane = threadIdx.x & 31;
    int fail = 0;
    for (int i = lane; i < n; i += 32) perm[i] = i;
    __syncwarp();
    for (int k = 0; k < n; ++k) {
        double best = -1.0; int bi = k;
        for (int i = k + lane; i < n; i += 32) { double v = fabs(A[i * ld + k]); if (v > best) { best = v; bi = i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(DSDF_FULL, best, o); int oi = __shfl_xor_sync(DSDF_FULL, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (!(best > 0.0)) fail = 1;
        if (bi != k) {
            for (int j = lane; j < n; j += 32) { double t = A[k * ld + j]; A[k * ld + j] = A[bi * ld + j]; A[bi * ld + j] = t; }
            if (lane == 0) { int t = perm[k]; perm[k] = perm[bi]; perm[bi] = t; }
        }
        __syncwarp();
        const double piv = A[k * ld + k];
        for (int i = k + 1 + lane; i < n; i += 32) {
            const double l = A[i * ld + k] / piv;
            A[i * ld + k] = l;
            for (int j = k + 1; j < n; ++j) A[i * ld + j] -= l * A[k * ld + j];
        }
        __syncwarp();
    }
    return fail;
}
// x <- (PLU)^-1 b ; b, x distinct shared vectors
DYN_FN void warp_lu_solve(const double* LU, int ld, int n, const int* perm, const double* b, double* x) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < n; i += 32) x[i] = b[perm[i]];
    __syncwarp();
    warp_trsv_lower_unit(LU, ld, n, x);
    warp_trsv_upper(LU, ld, n, x);
    __syncwarp();
}

// ---- load one world's problem into shared memory ---------------------------------------------------------------
DYN_FN void dyn_load(DynCtx& c, int w, const double* p, const double* v, const double* mass, const double* Ibody,
                         const double* fric, const double* rest, const double* f, double dtw, const int* count,
                         const int* cbody, const double* cgeo, const int* eq_rows, int maxc, int fd, double* mu_out,
                         double* e_out, double* h_out) {
    const int lane = threadIdx.x & 31;
    const DynSmem& L = c.L;
    const int nb = L.nb, nz = L.nz;
    c.nc = min(min(count[w], maxc), L.C);
    c.ni = c.nc * L.per;
    for (int i = lane; i < 2 * L.neq; i += 32) c.eq[i] = eq_rows[i];
    for (int i = lane; i < nz; i += 32) c.fpos[i] = 0;
    __syncwarp();
    for (int m = lane; m < L.neq; m += 32) c.fpos[6 * c.eq[2 * m] + c.eq[2 * m + 1]] = -(m + 1);
    __syncwarp();
    if (lane == 0) {
        int jf = 0;
        for (int i = 0; i < nz; ++i) if (c.fpos[i] == 0) { c.fidx[jf] = i; c.fpos[i] = jf++; }
    }
    __syncwarp();
    for (int b = lane; b < nb; b += 32) {
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
    for (int cc = lane; cc < c.nc; cc += 32) {
        const size_t oo = (size_t)w * maxc + cc;
        const int i1 = cbody[2 * oo], i2 = cbody[2 * oo + 1];
        c.cb[2 * cc] = i1; c.cb[2 * cc + 1] = i2;
        const double* g = cgeo + 10 * oo;
        const V3<double> n = v3<double>(g[0], g[1], g[2]), p1 = v3<double>(g[3], g[4], g[5]), p2 = v3<double>(g[6], g[7], g[8]);
        double* Gc = c.G + (size_t)cc * L.gs;
        row12<double>(p1, p2, n, Gc);
        V3<double> dirs[8];
        fdirs<double>(n, fd, dirs);
        for (int r = 0; r < L.half; ++r) row12<double>(p1, p2, dirs[r], Gc + 12 * (1 + r));   // rows r + half = -rows r
        mu_out[cc] = 0.5 * (fric[(size_t)w * nb + i1] + fric[(size_t)w * nb + i2]);
        e_out[cc] = (rest[(size_t)w * nb + i1] + rest[(size_t)w * nb + i2]) / 2;
    }
    __syncwarp();
    // u = M v + dt f
    double* pv = c.sv(DS_P);
    for (int i = lane; i < nz; i += 32) {
        const int b = i / 6, k = i % 6;
        const double* Q = c.Qb + 36 * b;
        double acc = 0.0;
        for (int j = 0; j < 6; ++j) acc += Q[6 * k + j] * v[(size_t)w * nz + 6 * b + j];
        pv[i] = acc + dtw * f[(size_t)w * nz + i];
    }
    // h = [e (Jc v); 0; 0]: one value per contact (normal row), zero on the friction and cone rows
    for (int cc = lane; cc < c.nc; cc += 32) {
        const double* Gc = c.G + (size_t)cc * L.gs;
        double jv = 0.0;
        for (int k = 0; k < 12; ++k) jv += Gc[k] * v[(size_t)w * nz + c.gidx(cc, k)];
        h_out[cc] = jv * e_out[cc];
    }
    __syncwarp();
}

// (F z)[r] for the engine's F in contact-major rows
__device__ __forceinline__ double Fz_row(const DynCtx& c, const double* z, const double* mu, int r, int fd) {
    const int per = c.L.per, cc = r / per, j = r % per;
    if (j == 0) return 0.0;
    if (j <= fd) return z[cc * per + per - 1];
    double acc = mu[cc] * z[cc * per];
    for (int q = 1; q <= fd; ++q) acc -= z[cc * per + q];
    return acc;
}
// out = G x for all rows (contact-major): friction row q + half is the negative of row q, the cone row is zero
DYN_FN void Gx_all(const DynCtx& c, const double* x, double* out) {
    const int lane = threadIdx.x & 31, per = c.L.per, half = c.L.half, rows = 1 + half;
    for (int e = lane; e < c.nc * rows; e += 32) {
        const int cc = e / rows, j = e % rows;
        const double* Gc = c.G + (size_t)cc * c.L.gs + 12 * j;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 12; ++k) acc += Gc[k] * x[c.gidx(cc, k)];
        out[cc * per + j] = acc;
        if (j >= 1) out[cc * per + j + half] = -acc;
        else out[cc * per + per - 1] = 0.0;
    }
    __syncwarp();
}
// (G' w)[I]
DYN_FN double Gtw_row(const DynCtx& c, const double* w, int I, int fd) {
    const int per = c.L.per, half = c.L.half, b = I / 6, k6 = I % 6;
    double acc = 0.0;
    for (int cc = 0; cc < c.nc; ++cc) {
        int k;
        if (c.cb[2 * cc] == b) k = k6; else if (c.cb[2 * cc + 1] == b) k = 6 + k6; else continue;
        const double* Gc = c.G + (size_t)cc * c.L.gs + k;
        const double* wc = w + cc * per;
        double a2 = Gc[0] * wc[0];
        for (int q = 1; q <= half; ++q) a2 += Gc[12 * q] * (wc[q] - wc[q + half]);
        acc += a2;
    }
    return acc;
}
// solve B_c u = t for every contact in place on t (contact-major), thread per contact
DYN_FN void block_solve(const DynCtx& c, const double* mu, double* t, int fd) {
    const int lane = threadIdx.x & 31, per = c.L.per;
    for (int cc = lane; cc < c.nc; cc += 32) {
        double* tc = t + cc * per;
        const double* ac = c.a + cc * per;
        const double un = tc[0] * ac[0];                 // ac = 1/a: reciprocals of B's diagonal (no fp64 divides here)
        double sum = 0.0;
        for (int q = 1; q <= fd; ++q) sum += tc[q] * ac[q];
        const double ug = (tc[per - 1] - mu[cc] * un + sum) * c.den[cc];
        tc[0] = un;
        for (int q = 1; q <= fd; ++q) tc[q] = (tc[q] - ug) * ac[q];
        tc[per - 1] = ug;
    }
    __syncwarp();
}

// factor: a = 1/d, den, YG, K = [[Q + sum G' B^-1 G, A'],[A,0]] -> LU.  Returns 1 on a singular K.
DYN_FN int dyn_factor(DynCtx& c, const double* d, const double* mu, int fd) {
    const int lane = threadIdx.x & 31;
    const DynSmem& L = c.L;
    const int per = L.per, nz = L.nz, nF = L.nF, ld = L.ldF;
    // a = 1/d is B's diagonal (batch.py:496); the kernel keeps 1/a and 1/den so the hot loops multiply
    for (int r = lane; r < c.ni; r += 32) c.a[r] = 1.0 / (1.0 / d[r]);
    __syncwarp();
    for (int cc = lane; cc < c.nc; cc += 32) {
        const double* ac = c.a + cc * per;
        double dn = 1.0 / d[cc * per + per - 1];
        for (int q = 1; q <= fd; ++q) dn += ac[q];
        c.den[cc] = 1.0 / dn;
    }
    __syncwarp();
    const int half = L.half;
    for (int e = lane; e < c.nc * half; e += 32) {
        const int cc = e / half, q = 1 + e % half;
        const double* ac = c.a + cc * per;
        c.AP[(size_t)2 * e] = ac[q] + ac[q + half];
        c.AP[(size_t)2 * e + 1] = ac[q] - ac[q + half];
    }
    __syncwarp();
    for (int e = lane; e < c.nc * 12; e += 32) {
        const int cc = e / 12, k = e % 12;
        const double* Gc = c.G + (size_t)cc * L.gs + k;
        const double* ap = c.AP + (size_t)2 * cc * half;
        double acc = -mu[cc] * Gc[0] * c.a[cc * per];
        for (int q = 1; q <= half; ++q) acc += Gc[12 * q] * ap[2 * (q - 1) + 1];
        c.YG[e] = acc * c.den[cc];
    }
    __syncwarp();
    // M~_FF = Q_FF + sum_c G_c' B_c^-1 G_c restricted to the free components; thread per entry
    for (int e = lane; e < nF * nF; e += 32) {
        const int If = e / nF, jf = e % nF, I = c.fidx[If], J = c.fidx[jf], bI = I / 6, bJ = J / 6;
        double acc = bI == bJ ? c.Qb[36 * bI + 6 * (I % 6) + (J % 6)] : 0.0;
        for (int cc = 0; cc < c.nc; ++cc) {
            const int i1 = c.cb[2 * cc], i2 = c.cb[2 * cc + 1];
            int ki, kj;
            if (i1 == bI) ki = I % 6; else if (i2 == bI) ki = 6 + I % 6; else continue;
            if (i1 == bJ) kj = J % 6; else if (i2 == bJ) kj = 6 + J % 6; else continue;
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
    __syncwarp();
    return warp_lu(c.K, ld, nF, c.perm);
}

// dx of [[M~, A'], [A, 0]] [dx; dy] = rhs with A a 0/1 selection of the pinned components and rhs_y = 0 (b = 0, x_P = 0):
// dx_P = rhs_y (= 0), dx_F = (M~_FF)^-1 rhs_F.  The multipliers dy follow from the pinned rows of the un-reduced
// system once dz is known: dy = -rx_P - (G' dz)_P - (Q dx)_P   (see dyn_solve).
__device__ inline void kkt_solve(DynCtx& c, const double* rhs, double* dxy) {
    const int lane = threadIdx.x & 31;
    const DynSmem& L = c.L;
    const int nz = L.nz, nF = L.nF, ld = L.ldF;
    double *tF = c.sv(DS_TF), *xF = c.sv(DS_XF);
    for (int jf = lane; jf < nF; jf += 32) tF[jf] = rhs[c.fidx[jf]];
    __syncwarp();
    warp_lu_solve(c.K, ld, nF, c.perm, tF, xF);
    for (int I = lane; I < nz; I += 32) { const int fp = c.fpos[I]; dxy[I] = fp >= 0 ? xF[fp] : rhs[nz + (-fp - 1)]; }
    __syncwarp();
}

// KKT solve (batch.py:380-410 semantics).  rxy = [rx; ry] (NULL = 0), rs (NULL = 0), rz (NULL = 0).
// Outputs dxy = [dx; dy], ds, dz.  Scratch: vec(DV_T), sv(DS_RHS).
DYN_FN void dyn_solve(DynCtx& c, const double* d, const double* mu, int fd, const double* rxy, const double* rs,
                          const double* rz, double* dxy, double* ds, double* dz) {
    const int lane = threadIdx.x & 31;
    const DynSmem& L = c.L;
    const int nz = L.nz, nq = L.nq;
    double* t = c.vec(DV_T);
    double* rhs = c.sv(DS_RHS);
    for (int r = lane; r < c.ni; r += 32) t[r] = (rz ? rz[r] : 0.0) - (rs ? rs[r] / d[r] : 0.0);
    __syncwarp();
    for (int r = lane; r < c.ni; r += 32) dz[r] = t[r];          // keep t (dz doubles as a second copy)
    __syncwarp();
    block_solve(c, mu, t, fd);                                     // t <- u = B^-1 t
    for (int I = lane; I < nq; I += 32) {
        double acc = rxy ? -rxy[I] : 0.0;
        if (I < nz) acc -= Gtw_row(c, t, I, fd);
        rhs[I] = acc;
    }
    __syncwarp();
    kkt_solve(c, rhs, dxy);
    double* gx = t;                                                                 // t is dead from here on
    Gx_all(c, dxy, gx);
    for (int r = lane; r < c.ni; r += 32) dz[r] = gx[r] + dz[r];                    // G dx + t
    __syncwarp();
    block_solve(c, mu, dz, fd);
    for (int m = lane; m < L.neq; m += 32) {                                       // multipliers of the pinned rows
        const int b = c.eq[2 * m], k = c.eq[2 * m + 1], I = 6 * b + k;
        double acc = (rxy ? -rxy[I] : 0.0) - Gtw_row(c, dz, I, fd);
        for (int j = 0; j < 6; ++j) acc -= c.Qb[36 * b + 6 * k + j] * dxy[6 * b + j];
        dxy[nz + m] = acc;
    }
    __syncwarp();
    for (int r = lane; r < c.ni; r += 32) ds[r] = ((rs ? -rs[r] : 0.0) - dz[r]) / d[r];
    __syncwarp();
}

DYN_FN double dyn_ratio_step(const DynCtx& c, const double* v, const double* dv) {
    const int lane = threadIdx.x & 31;
    double mx = -INFINITY;
    for (int r = lane; r < c.ni; r += 32) mx = fmax(mx, -v[r] / dv[r]);
    mx = warp_max(mx);
    const double repl = fmax(1.0, mx);
    double mn = INFINITY;
    for (int r = lane; r < c.ni; r += 32) mn = fmin(mn, dv[r] > 0.0 ? repl : -v[r] / dv[r]);
    return warp_min(mn);
}

__device__ inline DynCtx dyn_ctx(double* sm, const DynSmem& L) {
    DynCtx c;
    c.L = L;
    c.G = sm + L.oG; c.YG = sm + L.oYG; c.a = sm + L.oA; c.den = sm + L.oDen; c.AP = sm + L.oAP; c.V = sm + L.oV; c.K = sm + L.oK;
    c.Qb = sm + L.oQ; c.Sv = sm + L.oS;
    int* ib = reinterpret_cast<int*>(sm + L.oI);
    c.cb = ib; c.eq = ib + 2 * L.C; c.perm = c.eq + 2 * L.neq; c.fidx = c.perm + L.nq; c.fpos = c.fidx + L.nz;
    c.nc = c.ni = 0;
    return c;
}

__global__ void __launch_bounds__(32, 12)
dyn_forward_kernel(const double* __restrict__ p, const double* __restrict__ v, const double* __restrict__ mass,
                   const double* __restrict__ Ibody, const double* __restrict__ fric, const double* __restrict__ rest,
                   const double* __restrict__ f, const double* __restrict__ dt, const unsigned char* __restrict__ active,
                   const int* __restrict__ count, const int* __restrict__ cbody, const double* __restrict__ cgeo,
                   const int* __restrict__ eq_rows, int nb, int neq, int maxc, int C, int fd,
                   double eps, int not_improved_lim, int max_iter,
                   double* __restrict__ xo, double* __restrict__ nvo, double* __restrict__ nuo, double* __restrict__ lamo,
                   double* __restrict__ so, int* __restrict__ status_o, int* __restrict__ iters_o,
                   const int* __restrict__ vmap, int* __restrict__ ctrl, int cmin, int last_class) {
    extern __shared__ double smd[];
    // loop mode (ctrl != NULL, dsdf_steploop.cu): CTA w solves virtual world w = one attempt (its own dt[w]) of the real
    // world ws = vmap[w], whose state / parameters / contacts are read in place; outputs are indexed by w
    const int w = blockIdx.x, lane = threadIdx.x & 31;
    if (ctrl && (loop_idle(ctrl) || w >= ctrl[CT_NVIRT])) return;
    const int ws = vmap ? vmap[w] : w;
    const DynSmem L = dyn_layout(nb, neq, C, fd);
    const int nz = L.nz, nq = L.nq, per = L.per;
    if (active && !active[w]) {               // inactive world: velocity passes through
        if (nvo) for (int i = lane; i < nz; i += 32) nvo[(size_t)w * nz + i] = v[(size_t)ws * nz + i];
        return;
    }
    DynCtx c = dyn_ctx(smd, L);
    // contact-count classes: one launch sized for the typical count, one for the few worlds with many contacts (a world
    // that gave up halving keeps a penetrating state with dozens of contacts; sizing every CTA for it costs occupancy)
    if (count[ws] <= cmin || (count[ws] > C && !last_class)) return;
    if (count[ws] > C) {                      // more contacts than this launch's shared memory holds
        for (int i = lane; i < nz; i += 32) { xo[(size_t)w * nz + i] = NAN; if (nvo) nvo[(size_t)w * nz + i] = NAN; }
        if (lane == 0) {
            status_o[w] = DSDF_LCP_TOO_LARGE; if (iters_o) iters_o[w] = 0;
            if (ctrl) atomicOr(&ctrl[CT_ABORT], DSDF_STEP_DYN_SMEM);      // the host re-launches with a larger C
        }
        return;
    }
    __shared__ double s_mu[64], s_e[64], s_h[64];   // per contact: friction coefficient, restitution, h (C <= 64)
    const double* mu = s_mu;
    dyn_load(c, ws, p, v, mass, Ibody, fric, rest, f, dt[w], count, cbody, cgeo, eq_rows, maxc, fd, s_mu, s_e, s_h);
    auto hrow = [&](int r) { return (r % per == 0) ? s_h[r / per] : 0.0; };
    const int niCap = maxc * per;
    for (int r = lane; r < niCap; r += 32) { lamo[(size_t)w * niCap + r] = 0.0; so[(size_t)w * niCap + r] = 0.0; }
    const int ni = c.ni;
    double *s = c.vec(DV_S), *z = c.vec(DV_Z), *rz = c.vec(DV_RZ), *dsa = c.vec(DV_DSA),
           *dza = c.vec(DV_DZA), *ds = c.vec(DV_DS), *dz = c.vec(DV_DZ),
           *d = c.vec(DV_D);
    double *xy = c.sv(DS_XY), *rxy = c.sv(DS_RXY), *dxya = c.sv(DS_DXYA), *dxy = c.sv(DS_DXY), *bxy = c.sv(DS_BXY),
           *pv = c.sv(DS_P);
    int status = 0, iters = 0;
    // One loop for the initial point (it = -1; batch.py:84-110: d = 1, solve_kkt(p, 0, -h, -b), shift s, z >= 1) and the
    // interior-point iterations (batch.py:115-231), so that the factorisation and the KKT solve are instantiated ONCE:
    // the kernel is instruction-cache bound otherwise (29 k SASS instructions with the phases written out).
    bool have_best = false;
    double best_res = INFINITY, mu_gap = 0.0, sz = 0.0;
    int stalled = 0;
    for (int it = -1; it < max_iter; ++it) {
        double res = 0.0;
        if (it < 0) {
            for (int r = lane; r < ni; r += 32) { d[r] = 1.0; rz[r] = -hrow(r); }
            for (int i = lane; i < nq; i += 32) rxy[i] = i < nz ? pv[i] : 0.0;      // ry = -b = 0
        } else {
            // residuals (batch.py:117-131)
            double nrx = 0.0, nry = 0.0, nrz = 0.0;
            sz = 0.0;
            for (int I = lane; I < nq; I += 32) {
                double acc;
                if (I < nz) {
                    const int b = I / 6, k = I % 6;
                    const int fp = c.fpos[I];
                    const double ay = fp < 0 ? xy[nz + (-fp - 1)] : 0.0;
                    double qx = 0.0;
                    for (int j = 0; j < 6; ++j) qx += xy[6 * b + j] * c.Qb[36 * b + 6 * k + j];
                    acc = ay + Gtw_row(c, z, I, fd) + qx + pv[I];
                    nrx += acc * acc;
                } else {
                    const int m = I - nz;
                    acc = xy[6 * c.eq[2 * m] + c.eq[2 * m + 1]];          // A x - b, b = 0
                    nry += acc * acc;
                }
                rxy[I] = acc;
            }
            Gx_all(c, xy, c.vec(DV_T));
            for (int r = lane; r < ni; r += 32) {
                const double val_ = c.vec(DV_T)[r] + s[r] - hrow(r) - Fz_row(c, z, mu, r, fd);
                rz[r] = val_;
                nrz += val_ * val_;
                sz += s[r] * z[r];
            }
            nrx = wsum(nrx); nry = wsum(nry); nrz = wsum(nrz); sz = wsum(sz);
            mu_gap = fabs(sz / ni);
            res = (L.neq > 0 ? sqrt(nry) : 0.0) + sqrt(nrz) + sqrt(nrx) + ni * mu_gap;
            for (int r = lane; r < ni; r += 32) d[r] = z[r] / s[r];
        }
        __syncwarp();
        if (dyn_factor(c, d, mu, fd)) { status |= DSDF_LCP_FACTOR_FAIL; if (it >= 0) break; }
        if (it >= 0) {
            iters = it + 1;
            if (!have_best || res < best_res) {
                have_best = true; best_res = res; stalled = 0;
                for (int i = lane; i < nq; i += 32) bxy[i] = xy[i];
                for (int r = lane; r < ni; r += 32) {        // best (lam, s) go straight to the outputs (reference row order)
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
            // pass 0: initial solve (it < 0) or affine direction (rs = z); pass 1: corrector (rs in rz)
            const double* a_rxy = pass == 0 ? rxy : nullptr;
            const double* a_rs = pass == 0 ? (it < 0 ? nullptr : z) : rz;
            const double* a_rz = pass == 0 ? rz : nullptr;
            double* o_xy = pass == 0 ? (it < 0 ? xy : dxya) : dxy;
            double* o_s = pass == 0 ? (it < 0 ? s : dsa) : ds;
            double* o_z = pass == 0 ? (it < 0 ? z : dza) : dz;
            dyn_solve(c, d, mu, fd, a_rxy, a_rs, a_rz, o_xy, o_s, o_z);
            if (it < 0) {
                if (ni > 0) {                                   // batch.py:100-110
                    double ms = INFINITY, mz = INFINITY;
                    for (int r = lane; r < ni; r += 32) { ms = fmin(ms, s[r]); mz = fmin(mz, z[r]); }
                    ms = warp_min(ms); mz = warp_min(mz);
                    for (int r = lane; r < ni; r += 32) {
                        if (ms < 0.0) s[r] -= ms - 1.0;
                        if (mz < 0.0) z[r] -= mz - 1.0;
                    }
                    __syncwarp();
                }
                init_done = true;
                break;
            }
            if (pass == 0) {
                double alpha = fmin(fmin(dyn_ratio_step(c, z, dza), dyn_ratio_step(c, s, dsa)), 1.0);
                double t3 = 0.0;
                for (int r = lane; r < ni; r += 32) t3 += (s[r] + alpha * dsa[r]) * (z[r] + alpha * dza[r]);
                t3 = wsum(t3);
                const double r3 = t3 / sz, sig = r3 * r3 * r3;
                // corrector: rs = (-mu sig + dsa dza) / s  (stored in rz, which is free now)
                for (int r = lane; r < ni; r += 32) rz[r] = (-mu_gap * sig + dsa[r] * dza[r]) / s[r];
                __syncwarp();
            }
        }
        if (init_done) {
            if (ni == 0) {                                      // no contacts: the equality-constrained solve is the answer
                for (int i = lane; i < nq; i += 32) bxy[i] = xy[i];
                have_best = true;
                break;
            }
            continue;
        }
        for (int i = lane; i < nq; i += 32) dxy[i] += dxya[i];
        for (int r = lane; r < ni; r += 32) { ds[r] += dsa[r]; dz[r] += dza[r]; }
        __syncwarp();
        const double alpha = fmin(0.999 * fmin(dyn_ratio_step(c, z, dz), dyn_ratio_step(c, s, ds)), 1.0);
        for (int i = lane; i < nq; i += 32) xy[i] += alpha * dxy[i];
        for (int r = lane; r < ni; r += 32) { s[r] += alpha * ds[r]; z[r] += alpha * dz[r]; }
        __syncwarp();
    }
    if (ni > 0 && have_best && best_res > 1.0) status |= DSDF_LCP_INACCURATE;
    __syncwarp();
    for (int i = lane; i < nz; i += 32) {
        xo[(size_t)w * nz + i] = have_best ? bxy[i] : NAN;
        if (nvo) nvo[(size_t)w * nz + i] = have_best ? -bxy[i] : NAN;       // engines.py:81-82
    }
    for (int m = lane; m < L.neq; m += 32) nuo[(size_t)w * L.neq + m] = have_best ? bxy[nz + m] : NAN;
    if (lane == 0) { status_o[w] = status; if (iters_o) iters_o[w] = iters; }
}

// ---- backward: implicit differentiation at the solution, contracted straight onto the physical inputs ---------
__global__ void __launch_bounds__(32, 10)
dyn_backward_kernel(const double* __restrict__ p, const double* __restrict__ v, const double* __restrict__ mass,
                    const double* __restrict__ Ibody, const double* __restrict__ fric, const double* __restrict__ rest,
                    const double* __restrict__ f, const double* __restrict__ dt, const unsigned char* __restrict__ active,
                    const int* __restrict__ count, const int* __restrict__ cbody, const double* __restrict__ cgeo,
                    const int* __restrict__ eq_rows, int nb, int neq, int maxc, int C, int fd,
                    int stop_contact_grad, int stop_friction_grad,
                    const double* __restrict__ xs, const double* __restrict__ lams, const double* __restrict__ ss,
                    const double* __restrict__ gnv,
                    double* __restrict__ gp, double* __restrict__ gv, double* __restrict__ gmass, double* __restrict__ gI,
                    double* __restrict__ gfric, double* __restrict__ grest, double* __restrict__ gf,
                    double* __restrict__ gdt, double* __restrict__ ggeo, int cmin, int last_class) {
    extern __shared__ double smd[];
    const int w = blockIdx.x, lane = threadIdx.x & 31;
    const DynSmem L = dyn_layout(nb, neq, C, fd);
    DynCtx c = dyn_ctx(smd, L);
    const int nz = L.nz, nq = L.nq, per = L.per, niCap = maxc * per;
    const bool masked = active && !active[w];
    // contact-count classes (see dyn_forward_kernel): masked worlds are written by the first-class launch only
    if (masked ? cmin >= 0 : (count[w] <= cmin || (count[w] > C && !last_class))) return;
    const bool on = !masked && count[w] <= C;
    if (!on) {
        for (int i = lane; i < nb * 7; i += 32) gp[(size_t)w * nb * 7 + i] = 0.0;
        for (int i = lane; i < nz; i += 32) {               // inactive world: new_v = v
            gv[(size_t)w * nz + i] = masked ? gnv[(size_t)w * nz + i] : 0.0;
            gf[(size_t)w * nz + i] = 0.0;
        }
        for (int i = lane; i < nb; i += 32) {
            gmass[(size_t)w * nb + i] = 0.0; gfric[(size_t)w * nb + i] = 0.0; grest[(size_t)w * nb + i] = 0.0;
        }
        for (int i = lane; i < nb * 9; i += 32) gI[(size_t)w * nb * 9 + i] = 0.0;
        for (int i = lane; i < maxc * 10; i += 32) ggeo[(size_t)w * maxc * 10 + i] = 0.0;
        if (lane == 0) gdt[w] = 0.0;
        return;
    }
    __shared__ double s_mu[64], s_e[64], s_h[64];
    const double dtw = dt[w];
    dyn_load(c, w, p, v, mass, Ibody, fric, rest, f, dtw, count, cbody, cgeo, eq_rows, maxc, fd, s_mu, s_e, s_h);
    const int ni = c.ni, nc = c.nc;
    const double* vw = v + (size_t)w * nz;
    double *lam = c.vec(DV_Z), *d = c.vec(DV_D), *dl = c.vec(DV_DZ), *ds = c.vec(DV_DS);
    double *g = c.sv(DS_RXY), *dxy = c.sv(DS_DXY), *zh = c.sv(DS_XY);
    for (int i = lane; i < nq; i += 32) { g[i] = i < nz ? -gnv[(size_t)w * nz + i] : 0.0; zh[i] = i < nz ? xs[(size_t)w * nz + i] : 0.0; }
    for (int r = lane; r < ni; r += 32) {
        const int rr = ref_row(r / per, r % per, nc, fd);
        lam[r] = lams[(size_t)w * niCap + rr];
        d[r] = fmax(lam[r], 1e-8) / fmax(ss[(size_t)w * niCap + rr], 1e-8);
    }
    __syncwarp();
    dyn_factor(c, d, s_mu, fd);
    dyn_solve(c, d, s_mu, fd, g, nullptr, nullptr, dxy, ds, dl);      // dx = dxy[:nz], dlam = dl
    // ---- dp (= du), dh, and the outer-product gradients, contracted on the fly
    // gv = M' du + Jc' (e o dh),  dh = -dlam ; gf = dt du ; gdt = f . du
    for (int I = lane; I < nz; I += 32) {
        const int b = I / 6, k = I % 6;
        double acc = 0.0;
        for (int i = 0; i < 6; ++i) acc += c.Qb[36 * b + 6 * i + k] * dxy[6 * b + i];
        for (int cc = 0; cc < nc; ++cc) {
            int kk;
            if (c.cb[2 * cc] == b) kk = k; else if (c.cb[2 * cc + 1] == b) kk = 6 + k; else continue;
            acc += c.G[(size_t)cc * L.gs + kk] * s_e[cc] * (-dl[cc * per]);
        }
        gv[(size_t)w * nz + I] = acc;
        gf[(size_t)w * nz + I] = dtw * dxy[I];
    }
    {
        double acc = 0.0;
        for (int I = lane; I < nz; I += 32) acc += f[(size_t)w * nz + I] * dxy[I];
        acc = wsum(acc);
        if (lane == 0) gdt[w] = acc;
    }
    // per body: dM = dQ + du v' with dQ = 1/2 (dx z' + z dx')
    for (int b = lane; b < nb; b += 32) {
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
        for (int cc = 0; cc < nc; ++cc) {
            if (c.cb[2 * cc] != b && c.cb[2 * cc + 1] != b) continue;
            // dF[cone row, normal col] = dlam_cone * lam_normal
            gfr += 0.5 * dl[cc * per + per - 1] * lam[cc * per];
            double jv = 0.0;
            for (int k = 0; k < 12; ++k) jv += c.G[(size_t)cc * L.gs + k] * vw[c.gidx(cc, k)];
            gre += 0.5 * (-dl[cc * per]) * jv;
        }
        gfric[(size_t)w * nb + b] = gfr;
        grest[(size_t)w * nb + b] = gre;
    }
    // contact geometry: dG row = dlam_r z' + lam_r dx'  (+ dh e v' on normal rows)
    for (int t = lane; t < maxc * 10; t += 32) {
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
                    acc += (dlr * zh[I] + lr * dxy[I] + dh * s_e[cc] * vw[I]) * row[k].d;
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

size_t dsdf_dynamics_solve_smem_bytes(int nb, int neq, int ncontacts, int fric_dirs) {
    return dyn_layout(nb, neq, ncontacts, fric_dirs).bytes;
}

static int dyn_check(int W, int nb, int neq, int maxc, int* C, int fd, size_t* smem) {
    if (W <= 0 || nb <= 0 || neq < 0 || maxc <= 0 || (fd != 8 && fd != 4)) return -1;
    if (*C <= 0 || *C > maxc) *C = maxc;
    if (*C > 64) return -2;
    *smem = dyn_layout(nb, neq, *C, fd).bytes;
    if (*smem > 227 * 1024) return -2;
    return 0;
}

static size_t g_fwd_smem = 0, g_bwd_smem = 0;

int dsdf_dynamics_solve_loop(const double* p, const double* v, const double* mass, const double* Ibody,
                             const double* fric, const double* rest, const double* f, const double* dt,
                             const unsigned char* active, const int32_t* count, const int32_t* cbody, const double* cgeo,
                             const int32_t* eq_rows, int W, int nb, int neq, int maxc, int ncontacts_smem, int fric_dirs,
                             double eps, int not_improved_lim, int max_iter,
                             double* x, double* new_v, double* nu, double* lam, double* s, int32_t* status, int32_t* iters,
                             const int32_t* vmap, int32_t* ctrl, int count_min, int last_class, void* stream) {
    size_t smem;
    int C = ncontacts_smem;
    int rc = dyn_check(W, nb, neq, maxc, &C, fric_dirs, &smem);
    if (rc) return rc;
    cudaError_t e = ensure_smem(dyn_forward_kernel, smem, &g_fwd_smem);
    if (e != cudaSuccess) return (int)e;
    dyn_forward_kernel<<<W, 32, smem, (cudaStream_t)stream>>>(p, v, mass, Ibody, fric, rest, f, dt, active, count, cbody,
                                                              cgeo, eq_rows, nb, neq, maxc, C, fric_dirs, eps,
                                                              not_improved_lim, max_iter, x, new_v, nu, lam, s, status, iters,
                                                              vmap, ctrl, count_min, last_class);
    return (int)cudaGetLastError();
}

int dsdf_dynamics_solve(const double* p, const double* v, const double* mass, const double* Ibody,
                        const double* fric, const double* rest, const double* f, const double* dt,
                        const unsigned char* active, const int32_t* count, const int32_t* cbody, const double* cgeo,
                        const int32_t* eq_rows, int W, int nb, int neq, int maxc, int ncontacts_smem, int fric_dirs,
                        double eps, int not_improved_lim, int max_iter,
                        double* x, double* new_v, double* nu, double* lam, double* s, int32_t* status, int32_t* iters,
                        void* stream) {
    return dsdf_dynamics_solve_loop(p, v, mass, Ibody, fric, rest, f, dt, active, count, cbody, cgeo, eq_rows, W, nb, neq,
                                    maxc, ncontacts_smem, fric_dirs, eps, not_improved_lim, max_iter, x, new_v, nu, lam, s,
                                    status, iters, nullptr, nullptr, -1, 1, stream);
}

int dsdf_dynamics_solve_backward_loop(const double* p, const double* v, const double* mass, const double* Ibody,
                                      const double* fric, const double* rest, const double* f, const double* dt,
                                      const unsigned char* active, const int32_t* count, const int32_t* cbody,
                                      const double* cgeo, const int32_t* eq_rows, int W, int nb, int neq, int maxc,
                                      int ncontacts_smem, int fric_dirs, int stop_contact_grad, int stop_friction_grad,
                                      const double* x, const double* lam, const double* s, const double* g_new_v,
                                      double* gp, double* gv, double* gmass, double* gI, double* gfric, double* grest,
                                      double* gf, double* gdt, double* ggeo, int count_min, int last_class, void* stream) {
    size_t smem;
    int C = ncontacts_smem;
    int rc = dyn_check(W, nb, neq, maxc, &C, fric_dirs, &smem);
    if (rc) return rc;
    cudaError_t e = ensure_smem(dyn_backward_kernel, smem, &g_bwd_smem);
    if (e != cudaSuccess) return (int)e;
    dyn_backward_kernel<<<W, 32, smem, (cudaStream_t)stream>>>(p, v, mass, Ibody, fric, rest, f, dt, active, count, cbody,
                                                               cgeo, eq_rows, nb, neq, maxc, C, fric_dirs,
                                                               stop_contact_grad, stop_friction_grad, x, lam, s, g_new_v, gp,
                                                               gv, gmass, gI, gfric, grest, gf, gdt, ggeo, count_min,
                                                               last_class);
    return (int)cudaGetLastError();
}

int dsdf_dynamics_solve_backward(const double* p, const double* v, const double* mass, const double* Ibody,
                                 const double* fric, const double* rest, const double* f, const double* dt,
                                 const unsigned char* active, const int32_t* count, const int32_t* cbody,
                                 const double* cgeo, const int32_t* eq_rows, int W, int nb, int neq, int maxc,
                                 int ncontacts_smem, int fric_dirs, int stop_contact_grad, int stop_friction_grad,
                                 const double* x, const double* lam, const double* s, const double* g_new_v,
                                 double* gp, double* gv, double* gmass, double* gI, double* gfric, double* grest,
                                 double* gf, double* gdt, double* ggeo, void* stream) {
    return dsdf_dynamics_solve_backward_loop(p, v, mass, Ibody, fric, rest, f, dt, active, count, cbody, cgeo, eq_rows, W,
                                             nb, neq, maxc, ncontacts_smem, fric_dirs, stop_contact_grad,
                                             stop_friction_grad, x, lam, s, g_new_v, gp, gv, gmass, gI, gfric, grest, gf,
                                             gdt, ggeo, -1, 1, stream);
}

}  // extern "C"
