/*
 * dsdf_b200 -- C ABI of the B200-native DiffSDFSim stepping hot path.
 *
 * Every entry point is extern "C", takes plain device pointers / sizes / a CUDA
 * stream (as void*), never allocates or frees caller memory, keeps no global
 * mutable state and returns an int status (0 = launched OK, <0 = argument or
 * capacity error, >0 = CUDA error code).  Per-world solver/search conditions
 * are reported through int32 status arrays on the device, never by exceptions.
 *
 * Each function names the reference interface it replaces (paths relative to
 * the upstream repository).  All floating point is IEEE binary64 ("f64"): the
 * reference computes in torch.double (lcp_physics/physics/utils.py:59).
 *
 * Batch convention: the leading dimension of every array is the world index W;
 * arrays are dense row-major with the documented trailing shape.
 */
#ifndef DSDF_B200_H
#define DSDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSDF_VERSION 1

/* per-world LCP status bits (int32) */
#define DSDF_LCP_Q_SINGULAR   1   /* LU of Q hit a zero pivot            (batch.py:417-424) */
#define DSDF_LCP_Q_NOT_SPD    2   /* SPD check failed                    (lcp.py:109-113)   */
#define DSDF_LCP_FACTOR_FAIL  4   /* LU of R + D^-1 hit a zero pivot -> best iterate returned (batch.py:134-137) */
#define DSDF_LCP_INACCURATE   8   /* best residual > 1                   (batch.py:165,229) */

/* SDF kinds */
#define DSDF_SDF_BOX      0
#define DSDF_SDF_SPHERE   1
#define DSDF_SDF_CYLINDER 2
#define DSDF_SDF_GRID     3

int dsdf_version(void);

/* ------------------------------------------------------------------ LCP ----
 * Replaces LCPFunction(...).forward  (lcp_physics/lcp/lcp.py:48-153) with
 * pre_factor_kkt / forward / factor_kkt / solve_kkt / get_step
 * (lcp_physics/lcp/solvers/batch.py:413-479, 70-231, 485-520, 380-410, 234-237).
 *
 *   minimise 1/2 z'Qz + p'z  s.t.  G z + s = h + F lam (mixed LCP), A z = b
 *
 * Q (W,nz,nz)  p (W,nz)  G (W,nineq,nz)  h (W,nineq)  A (W,neq,nz)  b (W,neq)
 * F (W,nineq,nineq).  nineq_w (W) optional (NULL = all worlds use nineq): the
 * number of ACTIVE inequality rows of each world (rows/cols beyond it ignored).
 * Outputs: x (W,nz) nu (W,neq) lam (W,nineq) s (W,nineq) = best-residual
 * iterate; status (W) bit mask; iters (W) iterations run.
 * ws: workspace of dsdf_lcp_workspace_bytes(...) bytes.
 * Every reduction the reference takes over the whole batch tensor is taken per
 * world here (the reference engine only ever runs nBatch = 1).
 */
size_t dsdf_lcp_workspace_bytes(int W, int nz, int neq, int nineq);
size_t dsdf_lcp_smem_bytes(int nz, int neq, int nineq);   /* > 227 KiB: problem too large for this kernel */
int dsdf_lcp_forward(const double* Q, const double* p, const double* G, const double* h,
                     const double* A, const double* b, const double* F, const int32_t* nineq_w,
                     int W, int nz, int neq, int nineq,
                     double eps, int not_improved_lim, int max_iter, int check_spd,
                     double* x, double* nu, double* lam, double* s,
                     int32_t* status, int32_t* iters, void* ws, void* stream);

/* Replaces LCPFunctionFn.backward (lcp_physics/lcp/lcp.py:156-213).
 * gz (W,nz) = dL/dzhat.  Outputs dQ (W,nz,nz) dp (W,nz) dG (W,nineq,nz) dh (W,nineq)
 * dA (W,neq,nz) db (W,neq) dF (W,nineq,nineq); any output pointer may be NULL.
 */
int dsdf_lcp_backward(const double* Q, const double* G, const double* A, const double* F,
                      const int32_t* nineq_w, const double* x, const double* nu, const double* lam,
                      const double* s, const double* gz, int W, int nz, int neq, int nineq,
                      double* dQ, double* dp, double* dG, double* dh, double* dA, double* db, double* dF,
                      int32_t* status, void* ws, void* stream);

/* ------------------------------------------------------------ SDF query ----
 * Replaces SDF3D.query_sdfs(pts_loc, return_grads)  (sdf_physics/physics3d/bodies.py:721-760) with the
 * analytic SDFs (bodies.py:38-125) and the grid SDF (bodies.py:203-257; ev_sdf_utils.grid_interp).
 * shape (W,4) = [a,b,c,scale]: box dims/scale | sphere r/scale | cylinder r/scale,h/scale | grid: unused.
 * grid: W grids of res^3 doubles, world w at grid + w*grid_world_stride (stride 0 = shared grid).
 * pts (W,N,3) body-frame points -> sdf (W,N), dir (W,N,3) unit gradient (NULL/want_dir=0 to skip).
 * Outside the cube |p| <= scale: sdf = scale, dir = 0.
 */
int dsdf_sdf_query(int kind, const double* shape, const double* grid, int res, long long grid_world_stride,
                   const double* pts, int W, int N, int want_dir, double* sdf, double* dir, void* stream);
/* VJP of the above w.r.t. pts with torch-autograd conventions (DiffGridSDF.backward for grids, bodies.py:253-257):
 * gpts (W,N,3) = gsdf * d sdf/d pts + gdir . d dir/d pts ; gsdf / gdir may be NULL. */
int dsdf_sdf_query_backward(int kind, const double* shape, const double* grid, int res, long long grid_world_stride,
                            const double* pts, int W, int N, const double* gsdf, const double* gdir, double* gpts,
                            void* stream);

/* ----------------------------------------------------------- integrator ----
 * Replaces Body3D.move (sdf_physics/physics3d/bodies.py:488-496): q <- standardize(quat(expmap(w dt)) * q),
 * x <- x + v dt.  p (W,nb,7) = [qw,qx,qy,qz,x,y,z], v (W,nb,6) = [w,v], dt (W), active (W) uint8 or NULL;
 * inactive worlds copy p through.  The world-frame inertia update of set_p (bodies.py:509-511) is fused
 * into the dynamics kernel that consumes it.
 */
int dsdf_integrate(const double* p, const double* v, const double* dt, const unsigned char* active, int W, int nb,
                   double* p_out, void* stream);
/* gp (W,nb,7), gv (W,nb,6), gdt (W,nb) [per body; caller sums over bodies] from gp_out (W,nb,7). */
int dsdf_integrate_backward(const double* p, const double* v, const double* dt, const unsigned char* active, int W,
                            int nb, const double* gp_out, double* gp, double* gv, double* gdt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSDF_B200_H */
