// Element-wise operators of the hot path:
//   dsdf_sdf_query[_backward]   SDF3D.query_sdfs            sdf_physics/physics3d/bodies.py:721-760 (+38-125, 203-257)
//   dsdf_integrate[_backward]   Body3D.move / set_p (pose)   sdf_physics/physics3d/bodies.py:488-511
// Backward kernels use forward-mode duals of the same device code (dsdf_math.cuh), one seed per input scalar.
#include "dsdf_sdf.cuh"

namespace dsdf {

__device__ __forceinline__ SdfShape load_shape(int kind, const double* shape4, const double* grid, int res) {
    SdfShape s;
    s.kind = kind; s.a = shape4[0]; s.b = shape4[1]; s.c = shape4[2]; s.scale = shape4[3];
    s.grid = grid; s.res = res;
    return s;
}

__global__ void __launch_bounds__(256)
sdf_query_kernel(int kind, const double* __restrict__ shape, const double* __restrict__ grid, int res,
                 long long grid_stride, const double* __restrict__ pts, int N, int want_dir,
                 double* __restrict__ sdf, double* __restrict__ dir) {
    const int w = blockIdx.y;
    const SdfShape sh = load_shape(kind, shape + 4 * (size_t)w, grid ? grid + (size_t)w * grid_stride : nullptr, res);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const size_t o = (size_t)w * N + i;
        V3<double> p = v3<double>(pts[3 * o], pts[3 * o + 1], pts[3 * o + 2]);
        SdfOut<double> r = sdf_query<double>(sh, p, want_dir != 0);
        sdf[o] = r.d;
        if (want_dir) { dir[3 * o] = r.n.x; dir[3 * o + 1] = r.n.y; dir[3 * o + 2] = r.n.z; }
    }
}

__global__ void __launch_bounds__(256)
sdf_query_bwd_kernel(int kind, const double* __restrict__ shape, const double* __restrict__ grid, int res,
                     long long grid_stride, const double* __restrict__ pts, int N,
                     const double* __restrict__ gsdf, const double* __restrict__ gdir, double* __restrict__ gpts) {
    const int w = blockIdx.y;
    const SdfShape sh = load_shape(kind, shape + 4 * (size_t)w, grid ? grid + (size_t)w * grid_stride : nullptr, res);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const size_t o = (size_t)w * N + i;
        const double gs = gsdf ? gsdf[o] : 0.0;
        const double g0 = gdir ? gdir[3 * o] : 0.0, g1 = gdir ? gdir[3 * o + 1] : 0.0, g2 = gdir ? gdir[3 * o + 2] : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            V3<Dual> p = v3<Dual>(Dual(pts[3 * o], k == 0), Dual(pts[3 * o + 1], k == 1), Dual(pts[3 * o + 2], k == 2));
            SdfOut<Dual> r = sdf_query<Dual>(sh, p, gdir != nullptr);
            gpts[3 * o + k] = gs * r.d.d + g0 * r.n.x.d + g1 * r.n.y.d + g2 * r.n.z.d;
        }
    }
}

// ---- integrator: q <- standardize(quat(expmap(w dt)) (x) q), x <- x + v dt  (bodies.py:488-491) ----------------------
template <class S>
__device__ __forceinline__ void integrate_one(const S* p, const S* v, S dt, S* out) {
    M3<S> R = expmap<S>(v3<S>(v[0] * dt, v[1] * dt, v[2] * dt));
    Q4<S> dq = mat2q<S>(R);
    Q4<S> q = qmul<S>(dq, q4<S>(p[0], p[1], p[2], p[3]));
    out[0] = q.w; out[1] = q.x; out[2] = q.y; out[3] = q.z;
    out[4] = p[4] + v[3] * dt; out[5] = p[5] + v[4] * dt; out[6] = p[6] + v[5] * dt;
}

__global__ void __launch_bounds__(128)
integrate_kernel(const double* __restrict__ p, const double* __restrict__ v, const double* __restrict__ dt,
                 const unsigned char* __restrict__ active, int W, int nb, double* __restrict__ po) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * nb) return;
    const int w = i / nb;
    double pi[7], vi[6], o[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) pi[k] = p[(size_t)i * 7 + k];
    if (active && !active[w]) {
#pragma unroll
        for (int k = 0; k < 7; ++k) po[(size_t)i * 7 + k] = pi[k];
        return;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) vi[k] = v[(size_t)i * 6 + k];
    integrate_one<double>(pi, vi, dt[w], o);
#pragma unroll
    for (int k = 0; k < 7; ++k) po[(size_t)i * 7 + k] = o[k];
}

__global__ void __launch_bounds__(128)
integrate_bwd_kernel(const double* __restrict__ p, const double* __restrict__ v, const double* __restrict__ dt,
                     const unsigned char* __restrict__ active, int W, int nb, const double* __restrict__ gpo,
                     double* __restrict__ gp, double* __restrict__ gv, double* __restrict__ gdt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * nb) return;
    const int w = i / nb;
    double g[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) g[k] = gpo[(size_t)i * 7 + k];
    if (active && !active[w]) {
#pragma unroll
        for (int k = 0; k < 7; ++k) gp[(size_t)i * 7 + k] = g[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) gv[(size_t)i * 6 + k] = 0.0;
        gdt[i] = 0.0;
        return;
    }
    for (int seed = 0; seed < 14; ++seed) {
        Dual pi[7], vi[6], o[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) pi[k] = Dual(p[(size_t)i * 7 + k], seed == k ? 1.0 : 0.0);
#pragma unroll
        for (int k = 0; k < 6; ++k) vi[k] = Dual(v[(size_t)i * 6 + k], seed == 7 + k ? 1.0 : 0.0);
        Dual dti(dt[w], seed == 13 ? 1.0 : 0.0);
        integrate_one<Dual>(pi, vi, dti, o);
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) acc += g[k] * o[k].d;
        if (seed < 7) gp[(size_t)i * 7 + seed] = acc;
        else if (seed < 13) gv[(size_t)i * 6 + seed - 7] = acc;
        else gdt[i] = acc;
    }
}

}  // namespace dsdf

using namespace dsdf;

extern "C" {

int dsdf_sdf_query(int kind, const double* shape, const double* grid, int res, long long grid_world_stride,
                   const double* pts, int W, int N, int want_dir, double* sdf, double* dir, void* stream) {
    if (W <= 0 || N < 0 || kind < 0 || kind > 3 || (kind == DSDF_SDF_GRID && (!grid || res < 2))) return -1;
    if (N == 0) return 0;
    int bx = (N + 255) / 256;
    if (bx > 148 * 8) bx = 148 * 8;
    sdf_query_kernel<<<dim3(bx, W), 256, 0, (cudaStream_t)stream>>>(kind, shape, grid, res, grid_world_stride, pts, N,
                                                                    want_dir, sdf, dir);
    return (int)cudaGetLastError();
}

int dsdf_sdf_query_backward(int kind, const double* shape, const double* grid, int res, long long grid_world_stride,
                            const double* pts, int W, int N, const double* gsdf, const double* gdir, double* gpts,
                            void* stream) {
    if (W <= 0 || N < 0 || kind < 0 || kind > 3 || (kind == DSDF_SDF_GRID && (!grid || res < 2))) return -1;
    if (N == 0) return 0;
    int bx = (N + 255) / 256;
    if (bx > 148 * 8) bx = 148 * 8;
    sdf_query_bwd_kernel<<<dim3(bx, W), 256, 0, (cudaStream_t)stream>>>(kind, shape, grid, res, grid_world_stride, pts,
                                                                        N, gsdf, gdir, gpts);
    return (int)cudaGetLastError();
}

int dsdf_integrate(const double* p, const double* v, const double* dt, const unsigned char* active, int W, int nb,
                   double* p_out, void* stream) {
    if (W <= 0 || nb <= 0) return -1;
    integrate_kernel<<<(W * nb + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p, v, dt, active, W, nb, p_out);
    return (int)cudaGetLastError();
}

int dsdf_integrate_backward(const double* p, const double* v, const double* dt, const unsigned char* active, int W,
                            int nb, const double* gp_out, double* gp, double* gv, double* gdt, void* stream) {
    if (W <= 0 || nb <= 0) return -1;
    integrate_bwd_kernel<<<(W * nb + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p, v, dt, active, W, nb, gp_out, gp,
                                                                                 gv, gdt);
    return (int)cudaGetLastError();
}

}  // extern "C"
