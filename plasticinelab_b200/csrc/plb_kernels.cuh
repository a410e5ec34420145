// __global__ kernels (sm_100a): thin launch wrappers around the per-thread bodies in plb_bodies.cuh, plus the
// reductions (primitive pose gradients, loss terms) that need warp/block cooperation.
#pragma once
#include <cuda_runtime.h>
#include "plb_bodies.cuh"

namespace plb {

constexpr int kBlock = 128;

// poses of frame pf and pf+1 converted to T in shared memory (2 * n_prim * 8 doubles are read per block)
template <class T>
__device__ __forceinline__ void load_poses_smem(const double* __restrict__ traj, int pf, int n_prim, Pose<T>* s0, Pose<T>* s1) {
    if (threadIdx.x < 2 * n_prim) {
        int which = threadIdx.x / n_prim, k = threadIdx.x % n_prim;
        const double* src = traj + ((long long)(pf + which) * PLB_MAX_PRIM + k) * PLB_POSE_DIM;
        Pose<T> p = load_pose<T>(src);
        if (which == 0) s0[k] = p; else s1[k] = p;
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// warp-reduce one PoseGrad and add it to dst[8] (double) with 8 atomics from lane 0
template <class T>
__device__ __forceinline__ void reduce_pose_grad(const PoseGrad<T>& g, double* dst) {
    double vals[8] = {(double)g.pos.x, (double)g.pos.y, (double)g.pos.z, (double)g.rot.w, (double)g.rot.x,
                      (double)g.rot.y, (double)g.rot.z, (double)g.gap};
#pragma unroll
    for (int c = 0; c < 8; c++) {
        double s = warp_sum(vals[c]);
        if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(dst + c, s);
    }
}

// ------------------------------------------------------------------------------------------------ substep
template <class T>
__global__ void __launch_bounds__(kBlock) k_p2g(SimConst<T> P, T* frames, long long n_pad, int slot_in, int slot_out,
                                                int store_F_out, Material<T> mat, Vec4<T>* grid_in) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    p2g_body<T>(p, P, frame_at(frames, slot_in, n_pad), frame_at(frames, slot_out, n_pad), store_F_out != 0, mat, grid_in);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_grid_fwd(SimConst<T> P, PrimSet<T> prims, const double* traj, int pf,
                                                     Vec4<T>* grid_in, Vec4<T>* grid_out, int clear_in, long long n_nodes) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_nodes) return;
    grid_fwd_body<T>(node, P, prims, s0, s1, grid_in, grid_out, clear_in != 0);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_g2p(SimConst<T> P, T* frames, long long n_pad, int slot_in, int slot_out,
                                                const Vec4<T>* grid_out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    g2p_body<T>(p, P, frame_at(frames, slot_in, n_pad), frame_at(frames, slot_out, n_pad), grid_out);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_g2p_bwd(SimConst<T> P, T* frames, long long n_pad, int slot_in, T* adj_next,
                                                    T* adj_cur, const Vec4<T>* grid_out, Vec4<T>* g_out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    g2p_bwd_body<T>(p, P, frame_at(frames, slot_in, n_pad), frame_at(adj_next, 0, n_pad), frame_at(adj_cur, 0, n_pad),
                    grid_out, g_out);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_grid_bwd(SimConst<T> P, PrimSet<T> prims, const double* traj, int pf,
                                                     Vec4<T>* grid_in, Vec4<T>* g_out, Vec4<T>* g_in, int clear,
                                                     double* prim_grad, long long n_nodes) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PoseGrad<T> g0[PLB_MAX_PRIM], g1[PLB_MAX_PRIM];
    unsigned touched = 0;
    if (node < n_nodes) {
        for (int k = 0; k < P.n_prim; k++) { g0[k].clear(); g1[k].clear(); }
        grid_bwd_body<T>(node, P, prims, s0, s1, grid_in, g_out, g_in, clear != 0, g0, g1, touched);
    }
    for (int k = 0; k < P.n_prim; k++) {
        unsigned any = __ballot_sync(0xffffffffu, (touched >> k) & 1u);
        if (any == 0) continue;
        if (!((touched >> k) & 1u)) { g0[k].clear(); g1[k].clear(); }
        reduce_pose_grad<T>(g0[k], prim_grad + ((long long)pf * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
        reduce_pose_grad<T>(g1[k], prim_grad + ((long long)(pf + 1) * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
    }
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_p2g_bwd(SimConst<T> P, T* frames, long long n_pad, int slot_in, T* adj_next,
                                                    T* adj_cur, Material<T> mat, const Vec4<T>* g_in) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    p2g_bwd_body<T>(p, P, frame_at(frames, slot_in, n_pad), frame_at(adj_next, 0, n_pad), frame_at(adj_cur, 0, n_pad), mat, g_in);
}

// ------------------------------------------------------------------------------------------------ loss
// accumulator layout (doubles): [0] density  [1] sdf  [2] sum m*t  [3] sum m  [4] max m (bits)  [8+k] min_dist_k (bits)
constexpr int kAccN = 8 + PLB_MAX_PRIM;

template <class T>
__global__ void __launch_bounds__(kBlock) k_loss_mass(SimConst<T> P, T* frames, long long n_pad, int slot, T* grid_mass) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    loss_mass_body<T>(p, P, frame_at(frames, slot, n_pad), grid_mass);
}

__global__ void k_loss_init(double* acc) {
    int i = threadIdx.x;
    if (i < kAccN) acc[i] = (i >= 8) ? 100000.0 : 0.0;
}

template <class T>
__global__ void __launch_bounds__(256) k_loss_reduce(const T* __restrict__ grid_mass, const T* __restrict__ target,
                                                     const T* __restrict__ target_sdf, long long n_nodes, double* acc) {
    double d = 0, s = 0, mt = 0, sm = 0, mx = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += (long long)gridDim.x * blockDim.x) {
        double m = (double)grid_mass[i], t = (double)target[i];
        d += fabs(m - t);
        s += (double)target_sdf[i] * m;
        mt += m * t;
        sm += m;
        mx = fmax(mx, m);
    }
    d = warp_sum(d); s = warp_sum(s); mt = warp_sum(mt); sm = warp_sum(sm);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) {
        if (d != 0.0) atomicAdd(acc + 0, d);
        if (s != 0.0) atomicAdd(acc + 1, s);
        if (mt != 0.0) atomicAdd(acc + 2, mt);
        if (sm != 0.0) atomicAdd(acc + 3, sm);
        atomicMax(reinterpret_cast<unsigned long long*>(acc + 4), (unsigned long long)__double_as_longlong(mx));
    }
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_loss_contact(SimConst<T> P, PrimSet<T> prims, const double* traj, int pf,
                                                         T* frames, long long n_pad, int slot, double* acc) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = p < P.n_particles;
    V3<T> x = ok ? load_x(frame_at(frames, slot, n_pad), p) : zero3<T>();
    for (int k = 0; k < P.n_prim; k++) {
        if (!prims.s[k].movable) continue;
        double d = 1e30;
        if (ok) d = (double)tmax(tmax(prim_sdf(prims.s[k], s0[k], x), T(0)), T(0));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d = fmin(d, __shfl_xor_sync(0xffffffffu, d, o));
        if ((threadIdx.x & 31) == 0)
            atomicMin(reinterpret_cast<unsigned long long*>(acc + 8 + k), (unsigned long long)__double_as_longlong(d));
    }
}

// One thread: folds the accumulators into the step loss and the running loss; writes a record of 8 doubles.
template <class T>
__global__ void k_loss_finalize(PrimSet<T> prims, int n_prim, LossWeights w, const double* acc, double target_max,
                                double target_sum, double* loss_total, double* record) {
    double contact = 0;
    for (int k = 0; k < n_prim; k++)
        if (prims.s[k].movable) contact += acc[8 + k] * acc[8 + k];
    double step = contact * w.contact + acc[0] * w.density + acc[1] * w.sdf;
    *loss_total += step;
    double ma = __longlong_as_double((long long)reinterpret_cast<const unsigned long long*>(acc)[4]);
    double I = acc[2] / ma / target_max;
    double U = acc[3] / ma + target_sum / target_max;
    record[0] = *loss_total; record[1] = contact; record[2] = acc[0]; record[3] = acc[1];
    record[4] = I / (U - I); record[5] = step; record[6] = 0; record[7] = 0;
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_loss_bwd(SimConst<T> P, PrimSet<T> prims, const double* traj, int pf, T* frames,
                                                     long long n_pad, int slot, T* adj, const T* grid_mass, const T* target,
                                                     const T* target_sdf, LossWeights w, const double* acc, int contact_all,
                                                     double* prim_grad) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    PoseGrad<T> g[PLB_MAX_PRIM];
    unsigned touched = 0;
    if (p < P.n_particles) {
        for (int k = 0; k < P.n_prim; k++) g[k].clear();
        loss_bwd_body<T>(p, P, frame_at(frames, slot, n_pad), frame_at(adj, 0, n_pad), grid_mass, target, target_sdf,
                         (T)w.sdf, (T)w.density, (T)w.contact, prims, s0, acc + 8, contact_all, g, touched);
    }
    for (int k = 0; k < P.n_prim; k++) {
        unsigned any = __ballot_sync(0xffffffffu, (touched >> k) & 1u);
        if (any == 0) continue;
        if (!((touched >> k) & 1u)) g[k].clear();
        reduce_pose_grad<T>(g[k], prim_grad + ((long long)pf * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
    }
}

// ------------------------------------------------------------------------------------------------ target SDF
// One Jacobi sweep of Loss.update_target_sdf (plb/engine/losses/loss.py:81-101): every node looks at the 6^3-1
// neighbours with offsets in [-3,3)^3 in lexicographic order and keeps the nearest propagated surface point.
__global__ void __launch_bounds__(128) k_sdf_sweep(int n, double dx, const double* __restrict__ density,
                                                   const double* __restrict__ sdf_c, const double* __restrict__ near_c,
                                                   double* sdf_o, double* near_o, int* changed) {
    long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n * n * n;
    if (node >= total) return;
    int k = (int)(node % n), j = (int)((node / n) % n), i = (int)(node / ((long long)n * n));
    const double inf = 1000.0;
    double gx = i * dx, gy = j * dx, gz = k * dx;
    double best = inf;
    double bx = near_c[node * 3 + 0], by = near_c[node * 3 + 1], bz = near_c[node * 3 + 2];
    if (density[node] > 1e-4) {
        best = 0.0; bx = gx; by = gy; bz = gz;
    } else {
        for (int a = -3; a < 3; a++)
            for (int b = -3; b < 3; b++)
                for (int c = -3; c < 3; c++) {
                    int vi = i + a, vj = j + b, vk = k + c;
                    if (vi < 0 || vj < 0 || vk < 0 || vi >= n || vj >= n || vk >= n) continue;
                    if (a == 0 && b == 0 && c == 0) continue;
                    long long v = ((long long)vi * n + vj) * n + vk;
                    if (sdf_c[v] < inf) {
                        double px = near_c[v * 3 + 0], py = near_c[v * 3 + 1], pz = near_c[v * 3 + 2];
                        double ddx = gx - px, ddy = gy - py, ddz = gz - pz;
                        double dist = sqrt(ddx * ddx + ddy * ddy + ddz * ddz + 1e-8);
                        if (dist < best) { best = dist; bx = px; by = py; bz = pz; }
                    }
                }
    }
    if (best != sdf_c[node] || bx != near_c[node * 3] || by != near_c[node * 3 + 1] || bz != near_c[node * 3 + 2]) *changed = 1;
    sdf_o[node] = best;
    near_o[node * 3 + 0] = bx; near_o[node * 3 + 1] = by; near_o[node * 3 + 2] = bz;
}

// ------------------------------------------------------------------------------------------------ host <-> frame
// AoS float64 host layout (x[N][3], v[N][3], F[N][3][3], C[N][3][3], staged on the device) <-> packed planes
template <class T>
__global__ void k_pack_frame(int n, long long n_pad, T* frame, const double* x, const double* v, const double* F, const double* C) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    FramePtr<T> f = frame_at(frame, 0, n_pad);
    V3<T> xx, vv; M3<T> CC;
    load_xvC(f, p, xx, vv, CC);
    M3<T> FF = load_F(f, p);
    if (x) xx = mk3<T>((T)x[p * 3], (T)x[p * 3 + 1], (T)x[p * 3 + 2]);
    if (v) vv = mk3<T>((T)v[p * 3], (T)v[p * 3 + 1], (T)v[p * 3 + 2]);
    if (C) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) CC.m[i][j] = (T)C[p * 9 + i * 3 + j];
    if (F) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) FF.m[i][j] = (T)F[p * 9 + i * 3 + j];
    store_xvC(f, p, xx, vv, CC);
    store_F(f, p, FF);
}

template <class T>
__global__ void k_unpack_frame(int n, long long n_pad, T* frame, double* x, double* v, double* F, double* C) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    FramePtr<T> f = frame_at(frame, 0, n_pad);
    V3<T> xx, vv; M3<T> CC;
    load_xvC(f, p, xx, vv, CC);
    M3<T> FF = load_F(f, p);
    if (x) { x[p * 3] = xx.x; x[p * 3 + 1] = xx.y; x[p * 3 + 2] = xx.z; }
    if (v) { v[p * 3] = vv.x; v[p * 3 + 1] = vv.y; v[p * 3 + 2] = vv.z; }
    if (C) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[p * 9 + i * 3 + j] = CC.m[i][j];
    if (F) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[p * 9 + i * 3 + j] = FF.m[i][j];
}

template <class T> __global__ void k_convert(long long n, const double* src, T* dst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (T)src[i];
}
template <class T> __global__ void k_grid_to_double(long long n, const Vec4<T>* src, double* dst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { Vec4<T> v = src[i]; dst[i * 4] = v.x; dst[i * 4 + 1] = v.y; dst[i * 4 + 2] = v.z; dst[i * 4 + 3] = v.w; }
}
template <class T> __global__ void k_count_active(long long n, const Vec4<T>* grid_in, unsigned long long* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool a = (i < n) && (grid_in[i].w > T(1e-12));
    unsigned m = __ballot_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

}  // namespace plb
