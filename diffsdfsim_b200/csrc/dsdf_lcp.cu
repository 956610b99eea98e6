// Dense mixed-LCP primal-dual interior point, one CTA per world, fp64, KKT factorisation in shared memory.
//
// Replaces lcp_physics/lcp/lcp.py:48-213 + lcp_physics/lcp/solvers/batch.py:70-237,380-520 (see include/dsdf_b200.h).
// Linear algebra = the reference's block elimination of
//     S = [[A Q^-1 A', A Q^-1 G'], [G Q^-1 A', G Q^-1 G' + F + D^-1]]
// with pivoted LU of Q, of A Q^-1 A' and (per iteration) of T = R + D^-1 where
//     R = G Q^-1 G' + F - (G Q^-1 A')(A Q^-1 A')^-1 (A Q^-1 G')   (batch.py:435,474).
// R lives in a caller-provided global workspace (L2-resident), T in shared memory.
#include "dsdf_dense.cuh"
#include "../../include/dsdf_b200.h"

namespace dsdf {

struct LcpSmem {
    int nz, neq, ni;            // ni = capacity (max rows)
    int ldT, ldg, lde;          // leading dims (odd -> conflict-free column access)
    size_t off_T, off_G, off_A, off_Q, off_S11, off_S21, off_TT, off_vz, off_vi, off_ve, off_red, off_int;
    size_t bytes;
};

enum { NVZ = 8, NVI = 16, NVE = 8 };

__host__ __device__ inline LcpSmem lcp_layout(int nz, int neq, int ni) {
    LcpSmem L;
    L.nz = nz; L.neq = neq; L.ni = ni;
    L.ldT = ni | 1; L.ldg = nz | 1; L.lde = neq | 1;
    size_t o = 0;
    size_t tsz = (size_t)ni * L.ldT;
    size_t scratch = (size_t)nz * (ni + neq);          // Q^-1 G', Q^-1 A' during the one-time factorisation
    if (scratch < (size_t)nz * nz) scratch = (size_t)nz * nz;   // SPD-check copy of Q
    L.off_T = o;   o += tsz > scratch ? tsz : scratch;
    L.off_G = o;   o += (size_t)ni * L.ldg;
    L.off_A = o;   o += (size_t)neq * L.ldg;
    L.off_Q = o;   o += (size_t)nz * L.ldg;
    L.off_S11 = o; o += (size_t)neq * L.lde;
    L.off_S21 = o; o += (size_t)ni * L.lde;
    L.off_TT = o;  o += (size_t)neq * L.ldT;
    L.off_vz = o;  o += (size_t)NVZ * nz;
    L.off_vi = o;  o += (size_t)NVI * ni;
    L.off_ve = o;  o += (size_t)NVE * (neq > 0 ? neq : 1);
    L.off_red = o; o += 40;
    L.off_int = o; o += (size_t)(ni + nz + neq + 8 + 1) / 2 + 1;   // int32 area, counted in doubles
    L.bytes = o * sizeof(double);
    return L;
}

// All per-world solver state in shared memory.
struct LcpCtx {
    int nz, neq, ni;            // ni = ACTIVE rows of this world
    int ldT, ldg, lde;
    double *T, *G, *A, *Qlu, *S11, *S21, *TT, *red;
    double *vz, *vi, *ve;
    int *permT, *permQ, *permS, *ibuf;
    const double* Rg;           // global R (ld = ni capacity)
    int ldR;
    __device__ double* Z(int k) const { return vz + (size_t)k * nz; }
    __device__ double* I(int k) const { return vi + (size_t)k * niCap; }
    __device__ double* E(int k) const { return ve + (size_t)k * (neq > 0 ? neq : 1); }
    int niCap;
};

__device__ inline LcpCtx lcp_ctx(double* sm, const LcpSmem& L, int ni_active) {
    LcpCtx c;
    c.nz = L.nz; c.neq = L.neq; c.ni = ni_active; c.niCap = L.ni;
    c.ldT = L.ldT; c.ldg = L.ldg; c.lde = L.lde;
    c.T = sm + L.off_T; c.G = sm + L.off_G; c.A = sm + L.off_A; c.Qlu = sm + L.off_Q;
    c.S11 = sm + L.off_S11; c.S21 = sm + L.off_S21; c.TT = sm + L.off_TT; c.red = sm + L.off_red;
    c.vz = sm + L.off_vz; c.vi = sm + L.off_vi; c.ve = sm + L.off_ve;
    int* ib = reinterpret_cast<int*>(sm + L.off_int);
    c.permT = ib; c.permQ = ib + L.ni; c.permS = c.permQ + L.nz; c.ibuf = c.permS + L.neq;
    return c;
}

// ---- one-time factorisation (batch.py:413-479).  Returns status bits. Writes R to global. -------------
__device__ int lcp_prefactor(LcpCtx& c, const double* Qg, const double* Gg, const double* Ag, const double* Fg,
                             double* Rg, int ldR, int check_spd) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nz = c.nz, neq = c.neq, ni = c.ni;
    int status = 0;
    for (int e = tid; e < nz * nz; e += nt) c.Qlu[(e / nz) * c.ldg + e % nz] = Qg[e];
    for (int e = tid; e < ni * nz; e += nt) c.G[(e / nz) * c.ldg + e % nz] = Gg[e];
    for (int e = tid; e < neq * nz; e += nt) c.A[(e / nz) * c.ldg + e % nz] = Ag[e];
    if (tid == 0) { c.ibuf[0] = 0; c.ibuf[1] = 0; c.ibuf[2] = 0; }
    __syncthreads();
    if (check_spd) {
        // Cholesky attempt on a scratch copy (T region): fails iff Q (symmetric part read from the lower triangle) is not PD.
        double* C = c.T;
        for (int e = tid; e < nz * nz; e += nt) C[e] = c.Qlu[(e / nz) * c.ldg + e % nz];
        __syncthreads();
        if (tid < 32) {
            for (int k = 0; k < nz; ++k) {
                double dkk = C[k * nz + k];
                if (!(dkk > 0.0)) { if (tid == 0) c.ibuf[2] = 1; break; }
                double sq = sqrt(dkk);
                __syncwarp();
                for (int i = k + 1 + tid; i < nz; i += 32) C[i * nz + k] /= sq;
                __syncwarp();
                for (int i = k + 1 + tid; i < nz; i += 32)
                    for (int j = k + 1; j <= i; ++j) C[i * nz + j] -= C[i * nz + k] * C[j * nz + k];
                __syncwarp();
            }
        }
        __syncthreads();
        if (c.ibuf[2]) status |= DSDF_LCP_Q_NOT_SPD;
    }
    block_lu(c.Qlu, c.ldg, nz, c.permQ, &c.ibuf[0], &c.ibuf[1]);
    if (c.ibuf[1]) status |= DSDF_LCP_Q_SINGULAR;
    // Q^-1 G' (nz x ni) and Q^-1 A' (nz x neq), one thread per column, in the T region.
    double* QiGt = c.T;
    double* QiAt = c.T + (size_t)nz * ni;
    for (int j = tid; j < ni + neq; j += nt) {
        if (j < ni) thread_lu_solve(c.Qlu, c.ldg, nz, c.permQ, c.G + (size_t)j * c.ldg, 1, QiGt + j, ni);
        else        thread_lu_solve(c.Qlu, c.ldg, nz, c.permQ, c.A + (size_t)(j - ni) * c.ldg, 1, QiAt + (j - ni), neq);
    }
    __syncthreads();
    if (neq > 0) {
        for (int e = tid; e < neq * neq + ni * neq; e += nt) {
            const bool first = e < neq * neq;
            const int ee = first ? e : e - neq * neq;
            const int i = ee / neq, m = ee % neq;
            const double* row = first ? c.A + (size_t)i * c.ldg : c.G + (size_t)i * c.ldg;
            double acc = 0.0;
            for (int k = 0; k < nz; ++k) acc += row[k] * QiAt[k * neq + m];
            if (first) c.S11[i * c.lde + m] = acc; else c.S21[i * c.lde + m] = acc;
        }
        __syncthreads();
        if (tid == 0) c.ibuf[1] = 0;
        block_lu(c.S11, c.lde, neq, c.permS, &c.ibuf[0], &c.ibuf[1]);
        if (c.ibuf[1]) status |= DSDF_LCP_Q_SINGULAR;
        // TT = S11^-1 S21'  (neq x ni)
        for (int j = tid; j < ni; j += nt)
            thread_lu_solve(c.S11, c.lde, neq, c.permS, c.S21 + (size_t)j * c.lde, 1, c.TT + j, c.ldT);
        __syncthreads();
    }
    // R = G QiGt + F - S21 TT   -> global
    for (int e = tid; e < ni * ni; e += nt) {
        const int i = e / ni, j = e % ni;
        double acc = 0.0;
        const double* g = c.G + (size_t)i * c.ldg;
        for (int k = 0; k < nz; ++k) acc += g[k] * QiGt[k * ni + j];
        acc += Fg[(size_t)i * ldR + j];
        if (neq > 0) {
            double a2 = 0.0;
            const double* s21 = c.S21 + (size_t)i * c.lde;
            for (int m = 0; m < neq; ++m) a2 += s21[m] * c.TT[m * c.ldT + j];
            acc -= a2;
        }
        Rg[(size_t)i * ldR + j] = acc;
    }
    __syncthreads();
    c.Rg = Rg; c.ldR = ldR;
    return status;
}

// ---- factor T = R + diag(1/d)  (batch.py:485-520). Returns 1 on failure. -------------------------------
__device__ int lcp_factor(LcpCtx& c, const double* d) {
    const int tid = threadIdx.x, nt = blockDim.x, ni = c.ni;
    for (int e = tid; e < ni * ni; e += nt) {
        const int i = e / ni, j = e % ni;
        double v = c.Rg[(size_t)i * c.ldR + j];
        if (i == j) v += 1.0 / d[i];
        c.T[i * c.ldT + j] = v;
    }
    if (tid == 0) c.ibuf[1] = 0;
    __syncthreads();
    block_lu(c.T, c.ldT, ni, c.permT, &c.ibuf[0], &c.ibuf[1]);
    return c.ibuf[1];
}

// ---- KKT solve (batch.py:380-410).  rx(nz) rs(ni) rz(ni) ry(neq) -> dx ds dz dy.  rx/rs/rz/ry may be NULL = 0.
// Scratch: Z(6),Z(7) ; I(14),I(15) ; E(6),E(7).  Outputs must not alias inputs or scratch.
__device__ void lcp_solve(LcpCtx& c, const double* d, const double* rx, const double* rs, const double* rz,
                          const double* ry, double* dx, double* ds, double* dz, double* dy) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nz = c.nz, neq = c.neq, ni = c.ni;
    double *q = c.Z(6), *g1 = c.Z(7), *h2 = c.I(14), *t2 = c.I(15), *h1 = c.E(6), *y1 = c.E(7);
    if (rx) { block_lu_solve(c.Qlu, c.ldg, nz, c.permQ, rx, q); }
    else { for (int i = tid; i < nz; i += nt) q[i] = 0.0; __syncthreads(); }
    for (int i = tid; i < ni + neq; i += nt) {
        if (i < ni) {
            const double* g = c.G + (size_t)i * c.ldg;
            double acc = 0.0;
            for (int k = 0; k < nz; ++k) acc += g[k] * q[k];
            if (rs) acc += rs[i] / d[i];
            if (rz) acc -= rz[i];
            h2[i] = acc;
        } else {
            const int m = i - ni;
            const double* a = c.A + (size_t)m * c.ldg;
            double acc = 0.0;
            for (int k = 0; k < nz; ++k) acc += a[k] * q[k];
            if (ry) acc -= ry[m];
            h1[m] = acc;
        }
    }
    __syncthreads();
    if (neq > 0) {
        block_lu_solve(c.S11, c.lde, neq, c.permS, h1, y1);
        for (int i = tid; i < ni; i += nt) {
            const double* s21 = c.S21 + (size_t)i * c.lde;
            double acc = 0.0;
            for (int m = 0; m < neq; ++m) acc += s21[m] * y1[m];
            h2[i] -= acc;
        }
        __syncthreads();
    }
    block_lu_solve(c.T, c.ldT, ni, c.permT, h2, t2);
    for (int i = tid; i < ni; i += nt) {
        const double w2 = -t2[i];
        dz[i] = w2;
        ds[i] = ((rs ? -rs[i] : 0.0) - w2) / d[i];
    }
    __syncthreads();
    if (neq > 0) {
        for (int m = tid; m < neq; m += nt) {
            double acc = 0.0;
            const double* tt = c.TT + (size_t)m * c.ldT;
            for (int j = 0; j < ni; ++j) acc += tt[j] * dz[j];
            dy[m] = -y1[m] - acc;
        }
        __syncthreads();
    }
    for (int k = tid; k < nz; k += nt) {
        double acc = rx ? -rx[k] : 0.0;
        double a1 = 0.0;
        for (int i = 0; i < ni; ++i) a1 += c.G[(size_t)i * c.ldg + k] * dz[i];
        acc -= a1;
        if (neq > 0) {
            double a2 = 0.0;
            for (int m = 0; m < neq; ++m) a2 += c.A[(size_t)m * c.ldg + k] * dy[m];
            acc -= a2;
        }
        g1[k] = acc;
    }
    __syncthreads();
    block_lu_solve(c.Qlu, c.ldg, nz, c.permQ, g1, dx);
}

// get_step (batch.py:234-237): a = -v/dv; a[dv>0] = max(1, max a); min a.   IEEE semantics un-guarded.
__device__ double lcp_ratio_step(const LcpCtx& c, const double* v, const double* dv, double add_scale,
                                 const double* dv2) {
    // dv_eff = dv (+ dv2 if given)
    double mx = -INFINITY;
    for (int i = threadIdx.x; i < c.ni; i += blockDim.x) {
        double dd = dv[i] + (dv2 ? dv2[i] : 0.0);
        mx = fmax(mx, -v[i] / dd);
    }
    mx = block_reduce<RED_MAX>(mx, c.red);
    const double repl = fmax(1.0, mx);
    double mn = INFINITY;
    for (int i = threadIdx.x; i < c.ni; i += blockDim.x) {
        double dd = dv[i] + (dv2 ? dv2[i] : 0.0);
        double a = dd > 0.0 ? repl : -v[i] / dd;
        mn = fmin(mn, a);
    }
    (void)add_scale;
    return block_reduce<RED_MIN>(mn, c.red);
}

__global__ void __launch_bounds__(256)
lcp_forward_kernel(const double* __restrict__ Q, const double* __restrict__ p, const double* __restrict__ G,
                   const double* __restrict__ h, const double* __restrict__ A, const double* __restrict__ b,
                   const double* __restrict__ F, const int32_t* __restrict__ nineq_w,
                   int nz, int neq, int niCap, int niS, double eps, int not_improved_lim, int max_iter, int check_spd,
                   double* __restrict__ xo, double* __restrict__ nuo, double* __restrict__ lamo,
                   double* __restrict__ so, int32_t* __restrict__ status_o, int32_t* __restrict__ iters_o,
                   double* __restrict__ ws) {
    extern __shared__ double sm[];
    const int w = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const LcpSmem L = lcp_layout(nz, neq, niS);     // shared memory sized for niS rows; global strides use niCap
    if (nineq_w && nineq_w[w] < 0) return;          // masked world: outputs untouched
    if (nineq_w && nineq_w[w] > niS) {              // does not fit the launch's shared memory: flag, NaN outputs
        for (int i = tid; i < nz; i += nt) xo[(size_t)w * nz + i] = NAN;
        if (tid == 0) { status_o[w] = DSDF_LCP_TOO_LARGE; if (iters_o) iters_o[w] = 0; }
        return;
    }
    const int ni = nineq_w ? min(nineq_w[w], niCap) : niCap;
    LcpCtx c = lcp_ctx(sm, L, ni);
    const double* Qw = Q + (size_t)w * nz * nz;
    const double* Gw = G + (size_t)w * niCap * nz;
    const double* Aw = A ? A + (size_t)w * neq * nz : nullptr;
    const double* Fw = F + (size_t)w * niCap * niCap;
    double* Rw = ws + (size_t)w * niCap * niCap;
    // vectors
    double *x = c.Z(0), *pv = c.Z(1), *rx = c.Z(2), *dx = c.Z(3), *dxa = c.Z(4), *bx = c.Z(5);
    double *s = c.I(0), *z = c.I(1), *hv = c.I(2), *d = c.I(3), *rs = c.I(4), *rz = c.I(5), *ds = c.I(6), *dz = c.I(7),
           *dsa = c.I(8), *dza = c.I(9), *bz = c.I(10), *bs = c.I(11), *tmp = c.I(12);
    double *y = c.E(0), *ry = c.E(1), *dy = c.E(2), *dya = c.E(3), *by = c.E(4), *bv = c.E(5);

    int status = lcp_prefactor(c, Qw, Gw, Aw, Fw, Rw, niCap, check_spd);
    for (int i = tid; i < nz; i += nt) pv[i] = p[(size_t)w * nz + i];
    for (int i = tid; i < ni; i += nt) { hv[i] = h[(size_t)w * niCap + i]; d[i] = 1.0; tmp[i] = -hv[i]; }
    for (int i = tid; i < neq; i += nt) { bv[i] = b[(size_t)w * neq + i]; ry[i] = -bv[i]; }
    __syncthreads();

    int iters = 0;
    bool have_best = false;
    double best_res = INFINITY;
    if (ni > 0) {
        // initial point (batch.py:84-110): d = 1, solve_kkt(p, 0, -h, -b), shift s and z to >= 1
        int fail = lcp_factor(c, d);
        if (fail) status |= DSDF_LCP_FACTOR_FAIL;
        lcp_solve(c, d, pv, nullptr, tmp, neq > 0 ? ry : nullptr, x, s, z, y);
        __syncthreads();
        double ms = INFINITY, mz = INFINITY;
        for (int i = tid; i < ni; i += nt) { ms = fmin(ms, s[i]); mz = fmin(mz, z[i]); }
        ms = block_reduce<RED_MIN>(ms, c.red);
        mz = block_reduce<RED_MIN>(mz, c.red);
        for (int i = tid; i < ni; i += nt) {
            if (ms < 0.0) s[i] -= ms - 1.0;
            if (mz < 0.0) z[i] -= mz - 1.0;
        }
        __syncthreads();

        int stalled = 0;
        for (int it = 0; it < max_iter; ++it) {
            // residuals (batch.py:117-131)
            for (int k = tid; k < nz; k += nt) {
                double acc = 0.0;
                if (neq > 0) { double a = 0.0; for (int m = 0; m < neq; ++m) a += y[m] * c.A[(size_t)m * c.ldg + k]; acc = a; }
                double a1 = 0.0;
                for (int i = 0; i < ni; ++i) a1 += z[i] * c.G[(size_t)i * c.ldg + k];
                acc += a1;
                double a2 = 0.0;
                const double* qrow = Qw + (size_t)k * nz;
                for (int j = 0; j < nz; ++j) a2 += x[j] * qrow[j];      // x' Q'  -> row k of Q dotted with x
                acc += a2;
                rx[k] = acc + pv[k];
            }
            {   // rz = G x + s - h - F z ; warp per row for the F z part (coalesced global reads)
                const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
                for (int i = warp; i < ni; i += nw) {
                    const double* frow = Fw + (size_t)i * niCap;
                    double fz = 0.0;
                    for (int j = lane; j < ni; j += 32) fz += frow[j] * z[j];
                    fz = warp_sum(fz);
                    if (lane == 0) {
                        const double* g = c.G + (size_t)i * c.ldg;
                        double gx = 0.0;
                        for (int k = 0; k < nz; ++k) gx += x[k] * g[k];
                        rz[i] = gx + s[i] - hv[i] - fz;
                        rs[i] = z[i];
                    }
                }
            }
            for (int m = tid; m < neq; m += nt) {
                double acc = 0.0;
                const double* a = c.A + (size_t)m * c.ldg;
                for (int k = 0; k < nz; ++k) acc += x[k] * a[k];
                ry[m] = acc - bv[m];
            }
            __syncthreads();
            double sz = 0.0, nrz = 0.0, nrx = 0.0, nry = 0.0;
            for (int i = tid; i < ni; i += nt) { sz += s[i] * z[i]; nrz += rz[i] * rz[i]; }
            for (int i = tid; i < nz; i += nt) nrx += rx[i] * rx[i];
            for (int i = tid; i < neq; i += nt) nry += ry[i] * ry[i];
            sz = block_reduce<RED_SUM>(sz, c.red);
            nrz = block_reduce<RED_SUM>(nrz, c.red);
            nrx = block_reduce<RED_SUM>(nrx, c.red);
            nry = neq > 0 ? block_reduce<RED_SUM>(nry, c.red) : 0.0;
            const double mu = fabs(sz / ni);
            const double res = sqrt(nry) + sqrt(nrz) + sqrt(nrx) + ni * mu;
            for (int i = tid; i < ni; i += nt) d[i] = z[i] / s[i];
            __syncthreads();
            fail = lcp_factor(c, d);
            if (fail) { status |= DSDF_LCP_FACTOR_FAIL; break; }
            iters = it + 1;
            if (!have_best || res < best_res) {
                have_best = true; best_res = res; stalled = 0;
                for (int i = tid; i < nz; i += nt) bx[i] = x[i];
                for (int i = tid; i < ni; i += nt) { bz[i] = z[i]; bs[i] = s[i]; }
                for (int i = tid; i < neq; i += nt) by[i] = y[i];
            } else {
                ++stalled;
            }
            if (stalled == not_improved_lim || best_res < eps || mu > 1e32) break;

            // affine direction
            lcp_solve(c, d, rx, rs, rz, neq > 0 ? ry : nullptr, dxa, dsa, dza, dya);
            __syncthreads();
            double alpha = fmin(fmin(lcp_ratio_step(c, z, dza, 0, nullptr), lcp_ratio_step(c, s, dsa, 0, nullptr)), 1.0);
            double t3 = 0.0;
            for (int i = tid; i < ni; i += nt) t3 += (s[i] + alpha * dsa[i]) * (z[i] + alpha * dza[i]);
            t3 = block_reduce<RED_SUM>(t3, c.red);
            const double r3 = t3 / sz;
            const double sig = r3 * r3 * r3;
            for (int i = tid; i < ni; i += nt) tmp[i] = (-mu * sig + dsa[i] * dza[i]) / s[i];
            __syncthreads();
            // centering-corrector direction
            lcp_solve(c, d, nullptr, tmp, nullptr, nullptr, dx, ds, dz, dy);
            __syncthreads();
            for (int i = tid; i < nz; i += nt) dx[i] += dxa[i];
            for (int i = tid; i < ni; i += nt) { ds[i] += dsa[i]; dz[i] += dza[i]; }
            for (int i = tid; i < neq; i += nt) dy[i] += dya[i];
            __syncthreads();
            alpha = fmin(0.999 * fmin(lcp_ratio_step(c, z, dz, 0, nullptr), lcp_ratio_step(c, s, ds, 0, nullptr)), 1.0);
            for (int i = tid; i < nz; i += nt) x[i] += alpha * dx[i];
            for (int i = tid; i < ni; i += nt) { s[i] += alpha * ds[i]; z[i] += alpha * dz[i]; }
            for (int i = tid; i < neq; i += nt) y[i] += alpha * dy[i];
            __syncthreads();
        }
        if (have_best && best_res > 1.0) status |= DSDF_LCP_INACCURATE;
    } else {
        // no inequality rows: equality-constrained QP, same block elimination with an empty G
        // (engines.py:40-54 solves [[M,-Je'],[Je,0]] [v;l] = [u;0] by explicit inverse; same mathematics).
        lcp_solve(c, d, pv, nullptr, nullptr, neq > 0 ? ry : nullptr, bx, s, z, by);
        have_best = true;
    }
    __syncthreads();
    for (int i = tid; i < nz; i += nt) xo[(size_t)w * nz + i] = have_best ? bx[i] : NAN;
    for (int i = tid; i < niCap; i += nt) {
        const bool ok = have_best && i < ni;
        lamo[(size_t)w * niCap + i] = ok ? bz[i] : 0.0;
        so[(size_t)w * niCap + i] = ok ? bs[i] : 0.0;
    }
    for (int i = tid; i < neq; i += nt) nuo[(size_t)w * neq + i] = have_best ? by[i] : NAN;
    if (tid == 0) { status_o[w] = status; if (iters_o) iters_o[w] = iters; }
}

// ---- implicit backward (lcp.py:156-213) ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lcp_backward_kernel(const double* __restrict__ Q, const double* __restrict__ G, const double* __restrict__ A,
                    const double* __restrict__ F, const int32_t* __restrict__ nineq_w,
                    const double* __restrict__ xs, const double* __restrict__ nus, const double* __restrict__ lams,
                    const double* __restrict__ ss, const double* __restrict__ gz,
                    int nz, int neq, int niCap, int niS,
                    double* __restrict__ dQ, double* __restrict__ dp, double* __restrict__ dG, double* __restrict__ dh,
                    double* __restrict__ dA, double* __restrict__ db, double* __restrict__ dF,
                    int32_t* __restrict__ status_o, double* __restrict__ ws) {
    extern __shared__ double sm[];
    const int w = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const LcpSmem L = lcp_layout(nz, neq, niS);
    if (nineq_w && nineq_w[w] < 0) return;          // masked world: outputs untouched
    if (nineq_w && nineq_w[w] > niS) { if (tid == 0 && status_o) status_o[w] = DSDF_LCP_TOO_LARGE; return; }
    const int ni = nineq_w ? min(nineq_w[w], niCap) : niCap;
    LcpCtx c = lcp_ctx(sm, L, ni);
    const double* Qw = Q + (size_t)w * nz * nz;
    const double* Gw = G + (size_t)w * niCap * nz;
    const double* Aw = A ? A + (size_t)w * neq * nz : nullptr;
    const double* Fw = F + (size_t)w * niCap * niCap;
    double* Rw = ws + (size_t)w * niCap * niCap;
    double *g = c.Z(0), *dx = c.Z(1), *zh = c.Z(2);
    double *d = c.I(0), *ds = c.I(1), *dl = c.I(2), *lam = c.I(3);
    double *dnu = c.E(0), *nu = c.E(1);
    int status = lcp_prefactor(c, Qw, Gw, Aw, Fw, Rw, niCap, 0);
    for (int i = tid; i < nz; i += nt) { g[i] = gz[(size_t)w * nz + i]; zh[i] = xs[(size_t)w * nz + i]; }
    for (int i = tid; i < ni; i += nt) {
        lam[i] = lams[(size_t)w * niCap + i];
        d[i] = fmax(lam[i], 1e-8) / fmax(ss[(size_t)w * niCap + i], 1e-8);
    }
    for (int i = tid; i < neq; i += nt) nu[i] = nus[(size_t)w * neq + i];
    __syncthreads();
    if (ni > 0 && lcp_factor(c, d)) status |= DSDF_LCP_FACTOR_FAIL;
    lcp_solve(c, d, g, nullptr, nullptr, nullptr, dx, ds, dl, dnu);
    __syncthreads();
    if (dp) for (int i = tid; i < nz; i += nt) dp[(size_t)w * nz + i] = dx[i];
    if (dh) for (int i = tid; i < niCap; i += nt) dh[(size_t)w * niCap + i] = i < ni ? -dl[i] : 0.0;
    if (db) for (int i = tid; i < neq; i += nt) db[(size_t)w * neq + i] = -dnu[i];
    if (dQ) for (int e = tid; e < nz * nz; e += nt) {
        const int i = e / nz, j = e % nz;
        dQ[(size_t)w * nz * nz + e] = 0.5 * (dx[i] * zh[j] + zh[i] * dx[j]);
    }
    if (dG) for (int e = tid; e < niCap * nz; e += nt) {
        const int i = e / nz, j = e % nz;
        dG[(size_t)w * niCap * nz + e] = i < ni ? dl[i] * zh[j] + lam[i] * dx[j] : 0.0;
    }
    if (dA) for (int e = tid; e < neq * nz; e += nt) {
        const int i = e / nz, j = e % nz;
        dA[(size_t)w * neq * nz + e] = dnu[i] * zh[j] + nu[i] * dx[j];
    }
    if (dF) for (int e = tid; e < niCap * niCap; e += nt) {
        const int i = e / niCap, j = e % niCap;
        dF[(size_t)w * niCap * niCap + e] = (i < ni && j < ni) ? dl[i] * lam[j] : 0.0;
    }
    if (tid == 0 && status_o) status_o[w] = status;
}

}  // namespace dsdf

using namespace dsdf;

extern "C" {

int dsdf_version(void) { return DSDF_VERSION; }

size_t dsdf_lcp_workspace_bytes(int W, int nz, int neq, int nineq) {
    (void)nz; (void)neq;
    return (size_t)W * nineq * nineq * sizeof(double);
}

size_t dsdf_lcp_smem_bytes(int nz, int neq, int nineq) { return lcp_layout(nz, neq, nineq).bytes; }

static int lcp_check(int W, int nz, int neq, int nineq, int* nineq_smem, size_t* smem) {
    if (W <= 0 || nz <= 0 || neq < 0 || nineq < 0) return -1;
    if (*nineq_smem <= 0 || *nineq_smem > nineq) *nineq_smem = nineq;
    *smem = lcp_layout(nz, neq, *nineq_smem).bytes;
    if (*smem > 227 * 1024) return -2;
    return 0;
}

int dsdf_lcp_forward(const double* Q, const double* p, const double* G, const double* h,
                     const double* A, const double* b, const double* F, const int32_t* nineq_w,
                     int W, int nz, int neq, int nineq, int nineq_smem,
                     double eps, int not_improved_lim, int max_iter, int check_spd,
                     double* x, double* nu, double* lam, double* s,
                     int32_t* status, int32_t* iters, void* ws, void* stream) {
    size_t smem;
    int rc = lcp_check(W, nz, neq, nineq, &nineq_smem, &smem);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(lcp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lcp_forward_kernel<<<W, 256, smem, (cudaStream_t)stream>>>(Q, p, G, h, A, b, F, nineq_w, nz, neq, nineq, nineq_smem,
                                                               eps, not_improved_lim, max_iter, check_spd, x, nu, lam, s,
                                                               status, iters, (double*)ws);
    return (int)cudaGetLastError();
}

int dsdf_lcp_backward(const double* Q, const double* G, const double* A, const double* F,
                      const int32_t* nineq_w, const double* x, const double* nu, const double* lam,
                      const double* s, const double* gz, int W, int nz, int neq, int nineq, int nineq_smem,
                      double* dQ, double* dp, double* dG, double* dh, double* dA, double* db, double* dF,
                      int32_t* status, void* ws, void* stream) {
    size_t smem;
    int rc = lcp_check(W, nz, neq, nineq, &nineq_smem, &smem);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(lcp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lcp_backward_kernel<<<W, 256, smem, (cudaStream_t)stream>>>(Q, G, A, F, nineq_w, x, nu, lam, s, gz, nz, neq, nineq,
                                                                nineq_smem, dQ, dp, dG, dh, dA, db, dF, status,
                                                                (double*)ws);
    return (int)cudaGetLastError();
}

}  // extern "C"
