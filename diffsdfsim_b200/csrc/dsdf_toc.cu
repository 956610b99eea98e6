// Time-of-contact differential (World.H, lcp_physics/physics/world.py:141-237, and the gather of :275-327) for all
// worlds on the device.  The forward of H is the identity on dt, so only a backward kernel exists:
//   dL/dtheta += -(dD/dh)^+ dD/dtheta dL/dh          per world, over its NEW contacts (toc_mask)
// with the gap function of world.py:151-171
//   D(h) = n2 . (c2 - Rj(h)' (Ri(h) c1 + xi(h) - xj(h))),   Ri(h) = expmap(h w_i) R_i,  xi(h) = x_i + h v_i + a_i h^2 / 2
// evaluated at the start-of-step quantities the reference reconstructs from the end-of-step state (:275-327):
//   x_i = x_i' - h v_i,  R_i = expmap(-h w_i) R(q_i'),  c1 = R_1' p1,  c2 = R_2' p2,  n2 = R_2' n,  a = f / m.
// Derivatives: forward-mode duals of this composite, one seed per input scalar (42 per contact) + one for h.
// Conventions of H.backward kept: entries with dD/dh < TOL/h are zeroed (TOL = 1e-6, physics/utils.py:43), the
// pseudo-inverse denominator is guarded at 1e-5, and h also enters through the reconstruction (x_i, R_i).
#include "dsdf_math.cuh"
#include "dsdf_dense.cuh"
#include "../../include/dsdf_b200.h"

namespace dsdf {

enum { TOC_SEEDS = 42 };   // pp1 7 | pp2 7 | v1 6 | v2 6 | n 3 | p1 3 | p2 3 | a1 3 | a2 3 | hh 1   (a: linear part only)

template <class S> struct TocIn {
    S pp1[7], pp2[7], v1[6], v2[6], n[3], p1[3], p2[3], a1[3], a2[3], hh, hk;
};

template <class S> __device__ S toc_gap(const TocIn<S>& t) {
    const S hh = t.hh, hk = t.hk;
    // reconstruction of the start-of-step frame (world.py:296-318)
    M3<S> R1 = mat_mul(expmap<S>(v3<S>(-(hh * t.v1[0]), -(hh * t.v1[1]), -(hh * t.v1[2]))),
                       q2mat<S>(q4<S>(t.pp1[0], t.pp1[1], t.pp1[2], t.pp1[3])));
    M3<S> R2 = mat_mul(expmap<S>(v3<S>(-(hh * t.v2[0]), -(hh * t.v2[1]), -(hh * t.v2[2]))),
                       q2mat<S>(q4<S>(t.pp2[0], t.pp2[1], t.pp2[2], t.pp2[3])));
    V3<S> x1 = v3<S>(t.pp1[4] - hh * t.v1[3], t.pp1[5] - hh * t.v1[4], t.pp1[6] - hh * t.v1[5]);
    V3<S> x2 = v3<S>(t.pp2[4] - hh * t.v2[3], t.pp2[5] - hh * t.v2[4], t.pp2[6] - hh * t.v2[5]);
    V3<S> c1 = mat_applyT(R1, v3<S>(t.p1[0], t.p1[1], t.p1[2]));
    V3<S> c2 = mat_applyT(R2, v3<S>(t.p2[0], t.p2[1], t.p2[2]));
    V3<S> n2 = mat_applyT(R2, v3<S>(t.n[0], t.n[1], t.n[2]));
    // gap at time hk (world.py:151-171)
    M3<S> Rih = mat_mul(expmap<S>(v3<S>(hk * t.v1[0], hk * t.v1[1], hk * t.v1[2])), R1);
    M3<S> Rjh = mat_mul(expmap<S>(v3<S>(hk * t.v2[0], hk * t.v2[1], hk * t.v2[2])), R2);
    S half = cst(hk, 0.5);
    V3<S> xi = v3<S>(x1.x + hk * t.v1[3] + half * t.a1[0] * hk * hk, x1.y + hk * t.v1[4] + half * t.a1[1] * hk * hk,
                     x1.z + hk * t.v1[5] + half * t.a1[2] * hk * hk);
    V3<S> xj = v3<S>(x2.x + hk * t.v2[3] + half * t.a2[0] * hk * hk, x2.y + hk * t.v2[4] + half * t.a2[1] * hk * hk,
                     x2.z + hk * t.v2[5] + half * t.a2[2] * hk * hk);
    V3<S> ciw = mat_apply(Rih, c1) + xi;
    V3<S> cij = mat_applyT(Rjh, ciw - xj);
    return dot(n2, c2 - cij);
}

__device__ __forceinline__ void toc_load(TocIn<Dual>& t, int seed, const double* pp1, const double* pp2, const double* v1,
                                         const double* v2, const double* g, const double* a1, const double* a2, double h) {
    int s = 0;
    for (int k = 0; k < 7; ++k, ++s) t.pp1[k] = Dual(pp1[k], seed == s ? 1.0 : 0.0);
    for (int k = 0; k < 7; ++k, ++s) t.pp2[k] = Dual(pp2[k], seed == s ? 1.0 : 0.0);
    for (int k = 0; k < 6; ++k, ++s) t.v1[k] = Dual(v1[k], seed == s ? 1.0 : 0.0);
    for (int k = 0; k < 6; ++k, ++s) t.v2[k] = Dual(v2[k], seed == s ? 1.0 : 0.0);
    for (int k = 0; k < 3; ++k, ++s) t.n[k] = Dual(g[k], seed == s ? 1.0 : 0.0);
    for (int k = 0; k < 3; ++k, ++s) t.p1[k] = Dual(g[3 + k], seed == s ? 1.0 : 0.0);
    for (int k = 0; k < 3; ++k, ++s) t.p2[k] = Dual(g[6 + k], seed == s ? 1.0 : 0.0);
    for (int k = 0; k < 3; ++k, ++s) t.a1[k] = Dual(a1[k], seed == s ? 1.0 : 0.0);
    for (int k = 0; k < 3; ++k, ++s) t.a2[k] = Dual(a2[k], seed == s ? 1.0 : 0.0);
    t.hh = Dual(h, seed == s ? 1.0 : 0.0);      // s == 41
    t.hk = Dual(h, seed == TOC_SEEDS ? 1.0 : 0.0);
}

// One CTA per world.  part[c][seed] = w_c dD_c/dtheta_seed in shared memory, then thread-per-output sums over the
// contacts in index order (deterministic).
__global__ void __launch_bounds__(128)
toc_backward_kernel(int nb, int maxc, const double* __restrict__ dt, const unsigned char* __restrict__ toc_mask,
                    const int* __restrict__ cbody, const double* __restrict__ p, const double* __restrict__ v,
                    const double* __restrict__ geo, const double* __restrict__ f, const double* __restrict__ mass,
                    const double* __restrict__ g_dth, double base_tol, double* __restrict__ g_dt, double* __restrict__ gp,
                    double* __restrict__ gv, double* __restrict__ ggeo, double* __restrict__ gf,
                    double* __restrict__ gmass) {
    extern __shared__ double sm[];
    double* part = sm;                               // [maxc][TOC_SEEDS]
    double* dDh = sm + (size_t)maxc * TOC_SEEDS;     // [maxc] masked dD/dh, then the weights w_c
    __shared__ int s_any;
    const int w = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double h = dt[w], gh = g_dth[w];
    if (tid == 0) s_any = 0;
    __syncthreads();
    for (int c = tid; c < maxc; c += nt) if (toc_mask[(size_t)w * maxc + c]) s_any = 1;
    __syncthreads();
    const bool any = s_any != 0;
    // outputs default to zero / identity
    for (int i = tid; i < nb * 7; i += nt) gp[(size_t)w * nb * 7 + i] = 0.0;
    for (int i = tid; i < nb * 6; i += nt) { gv[(size_t)w * nb * 6 + i] = 0.0; gf[(size_t)w * nb * 6 + i] = 0.0; }
    for (int i = tid; i < nb; i += nt) gmass[(size_t)w * nb + i] = 0.0;
    for (int i = tid; i < maxc * 10; i += nt) ggeo[(size_t)w * maxc * 10 + i] = 0.0;
    if (tid == 0) g_dt[w] = gh;                      // H is the identity in dt (world.py:215)
    if (!any) return;
    __syncthreads();
    auto body_ptrs = [&](int c, const double*& pp1, const double*& pp2, const double*& v1, const double*& v2, double* a1,
                         double* a2, int& i1, int& i2) {
        const size_t oo = (size_t)w * maxc + c;
        i1 = min(max(cbody[2 * oo], 0), nb - 1); i2 = min(max(cbody[2 * oo + 1], 0), nb - 1);
        pp1 = p + ((size_t)w * nb + i1) * 7; pp2 = p + ((size_t)w * nb + i2) * 7;
        v1 = v + ((size_t)w * nb + i1) * 6; v2 = v + ((size_t)w * nb + i2) * 6;
        const double m1 = mass[(size_t)w * nb + i1], m2 = mass[(size_t)w * nb + i2];
        for (int k = 0; k < 3; ++k) {
            a1[k] = f[((size_t)w * nb + i1) * 6 + 3 + k] / m1;
            a2[k] = f[((size_t)w * nb + i2) * 6 + 3 + k] / m2;
        }
    };
    // pass A: dD/dh per contact
    for (int c = tid; c < maxc; c += nt) {
        double d = 0.0;
        if (toc_mask[(size_t)w * maxc + c]) {
            const double *pp1, *pp2, *v1, *v2; double a1[3], a2[3]; int i1, i2;
            body_ptrs(c, pp1, pp2, v1, v2, a1, a2, i1, i2);
            TocIn<Dual> t;
            toc_load(t, TOC_SEEDS, pp1, pp2, v1, v2, geo + ((size_t)w * maxc + c) * 10, a1, a2, h);
            d = toc_gap<Dual>(t).d;
            if (d < base_tol / h) d = 0.0;           // world.py:204
        }
        dDh[c] = d;
    }
    __syncthreads();
    double den = 0.0;
    for (int c = 0; c < maxc; ++c) den += dDh[c] * dDh[c];      // every thread: same order, same value
    __syncthreads();
    for (int c = tid; c < maxc; c += nt) dDh[c] = den > 1e-5 ? -(dDh[c] / den) * gh : 0.0;   // w_c (world.py:206-214)
    __syncthreads();
    // pass B: weighted input derivatives
    for (int item = tid; item < maxc * TOC_SEEDS; item += nt) {
        const int c = item / TOC_SEEDS, seed = item % TOC_SEEDS;
        double val_ = 0.0;
        if (toc_mask[(size_t)w * maxc + c] && dDh[c] != 0.0) {
            const double *pp1, *pp2, *v1, *v2; double a1[3], a2[3]; int i1, i2;
            body_ptrs(c, pp1, pp2, v1, v2, a1, a2, i1, i2);
            TocIn<Dual> t;
            toc_load(t, seed, pp1, pp2, v1, v2, geo + ((size_t)w * maxc + c) * 10, a1, a2, h);
            val_ = dDh[c] * toc_gap<Dual>(t).d;
        }
        part[item] = val_;
    }
    __syncthreads();
    // scatter-free accumulation: thread per output scalar, contacts in index order
    for (int o = tid; o < nb * 13 + nb * 4; o += nt) {
        if (o < nb * 13) {                           // pose (7) and velocity (6) of body b
            const int b = o / 13, k = o % 13;
            double acc = 0.0;
            for (int c = 0; c < maxc; ++c) {
                if (!toc_mask[(size_t)w * maxc + c]) continue;
                const size_t oo = (size_t)w * maxc + c;
                const int i1 = min(max(cbody[2 * oo], 0), nb - 1), i2 = min(max(cbody[2 * oo + 1], 0), nb - 1);
                const double* pc = part + (size_t)c * TOC_SEEDS;
                if (i1 == b) acc += k < 7 ? pc[k] : pc[14 + (k - 7)];
                if (i2 == b) acc += k < 7 ? pc[7 + k] : pc[20 + (k - 7)];
            }
            if (k < 7) gp[((size_t)w * nb + b) * 7 + k] = acc;
            else gv[((size_t)w * nb + b) * 6 + (k - 7)] = acc;
        } else {                                     // force (linear part) and mass of body b through a = f / m
            const int q = o - nb * 13, b = q / 4, k = q % 4;
            double ga[3] = {0.0, 0.0, 0.0};
            for (int c = 0; c < maxc; ++c) {
                if (!toc_mask[(size_t)w * maxc + c]) continue;
                const size_t oo = (size_t)w * maxc + c;
                const int i1 = min(max(cbody[2 * oo], 0), nb - 1), i2 = min(max(cbody[2 * oo + 1], 0), nb - 1);
                const double* pc = part + (size_t)c * TOC_SEEDS;
                for (int e = 0; e < 3; ++e) {
                    if (i1 == b) ga[e] += pc[35 + e];
                    if (i2 == b) ga[e] += pc[38 + e];
                }
            }
            const double m = mass[(size_t)w * nb + b];
            if (k < 3) gf[((size_t)w * nb + b) * 6 + 3 + k] = ga[k] / m;
            else {
                double acc = 0.0;
                for (int e = 0; e < 3; ++e) acc -= ga[e] * f[((size_t)w * nb + b) * 6 + 3 + e] / (m * m);
                gmass[(size_t)w * nb + b] = acc;
            }
        }
    }
    for (int o = tid; o < maxc * 9; o += nt) {       // contact tuple n, p1, p2 (pen does not enter)
        const int c = o / 9, k = o % 9;
        if (toc_mask[(size_t)w * maxc + c]) ggeo[((size_t)w * maxc + c) * 10 + k] = part[(size_t)c * TOC_SEEDS + 26 + k];
    }
    if (tid == 0) {
        double acc = gh;
        for (int c = 0; c < maxc; ++c) if (toc_mask[(size_t)w * maxc + c]) acc += part[(size_t)c * TOC_SEEDS + 41];
        g_dt[w] = acc;
    }
}

}  // namespace dsdf

extern "C" int dsdf_toc_backward(int W, int nb, int maxc, const double* dt, const unsigned char* toc_mask,
                                 const int32_t* cbody, const double* p, const double* v, const double* geo,
                                 const double* f, const double* mass, const double* g_dt_h, double base_tol,
                                 double* g_dt, double* gp, double* gv, double* ggeo, double* gf, double* gmass,
                                 void* stream) {
    if (W <= 0 || nb <= 0 || maxc <= 0) return -1;
    const size_t smem = ((size_t)maxc * dsdf::TOC_SEEDS + maxc) * sizeof(double);
    if (smem > 200 * 1024) return -2;
    cudaError_t e = cudaFuncSetAttribute(dsdf::toc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dsdf::toc_backward_kernel<<<W, 128, smem, (cudaStream_t)stream>>>(nb, maxc, dt, toc_mask, cbody, p, v, geo, f, mass,
                                                                      g_dt_h, base_tol, g_dt, gp, gv, ggeo, gf, gmass);
    return (int)cudaGetLastError();
}
