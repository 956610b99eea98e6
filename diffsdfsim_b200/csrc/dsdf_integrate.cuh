// Semi-implicit pose update of one body, shared by the integrator kernels (dsdf_pointops.cu) and the step loop
// (dsdf_steploop.cu).  Scalar-generic (double | Dual), see dsdf_math.cuh.
#pragma once
#include "dsdf_math.cuh"

namespace dsdf {

// ---- integrator: q <- standardize(quat(expmap(w dt)) (x) q), x <- x + v dt  (bodies.py:488-491) ----------------------
template <class S>
__device__ __forceinline__ void integrate_one(const S* p, const S* v, S dt, S* out) {
    M3<S> R = expmap<S>(v3<S>(v[0] * dt, v[1] * dt, v[2] * dt));
    Q4<S> dq = mat2q<S>(R);
    Q4<S> q = qmul<S>(dq, q4<S>(p[0], p[1], p[2], p[3]));
    out[0] = q.w; out[1] = q.x; out[2] = q.y; out[3] = q.z;
    out[4] = p[4] + v[3] * dt; out[5] = p[5] + v[4] * dt; out[6] = p[6] + v[5] * dt;
}

}  // namespace dsdf
