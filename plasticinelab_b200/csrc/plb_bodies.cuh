// Per-thread bodies of the simulation kernels ("one particle" / "one node"), shared by
//   * the __global__ kernels in plb_kernels.cu (sm_100a), and
//   * tests/host/emul.cpp, which runs them sequentially on the CPU for parity against the oracle.
//
// HBM layout of one particle frame (n = padded particle count, scalar type T, 24 n scalars, exact packing):
//   group A (written by G2P):  A0[n] = (x0,x1,x2,v0)  A1[n] = (v1,v2,C00,C01)  A2[n] = (C02,C10,C11,C12)
//                              a3[n] = C20   a4[n] = C21   a5[n] = C22
//   group B (written by P2G):  B0[n] = (F00,F01,F02,F10)  B1[n] = (F11,F12,F20,F21)  b2[n] = F22
// A warp reads each Vec4 plane as one fully coalesced 512 B (float) request and each scalar plane as 128 B.
// Adjoint frames use the same layout.  Grids are Vec4 per node: grid_in = (momentum xyz, mass),
// grid_out = (velocity xyz, 0) and likewise for their adjoints, linear index (i * n + j) * n + k.
#pragma once
#include "plb_grid.cuh"
#if defined(PLB_WARP_EMUL) && !defined(__CUDACC__)
#include <mutex>          // host warp emulation (tests/host/warp_emul.cpp): 32 threads per warp scatter concurrently
#endif

namespace plb {

template <class T> struct FramePtr {
    Vec4<T>* A0; Vec4<T>* A1; Vec4<T>* A2; T* a3; T* a4; T* a5;
    Vec4<T>* B0; Vec4<T>* B1; T* b2;
};

template <class T> PLB_HD FramePtr<T> frame_at(T* base, long long f, long long n_pad) {
    T* b = base + f * 24 * n_pad;
    FramePtr<T> r;
    r.A0 = reinterpret_cast<Vec4<T>*>(b);
    r.A1 = reinterpret_cast<Vec4<T>*>(b + 4 * n_pad);
    r.A2 = reinterpret_cast<Vec4<T>*>(b + 8 * n_pad);
    r.a3 = b + 12 * n_pad; r.a4 = b + 13 * n_pad; r.a5 = b + 14 * n_pad;
    r.B0 = reinterpret_cast<Vec4<T>*>(b + 15 * n_pad);
    r.B1 = reinterpret_cast<Vec4<T>*>(b + 19 * n_pad);
    r.b2 = b + 23 * n_pad;
    return r;
}

template <class T> struct Material { const T* mu; const T* lam; const T* ys; };   // null => uniform from SimConst

// ---- plane access
template <class T> PLB_HD void load_xvC(const FramePtr<T>& f, int p, V3<T>& x, V3<T>& v, M3<T>& C) {
    Vec4<T> q0 = f.A0[p], q1 = f.A1[p], q2 = f.A2[p];
    x = mk3<T>(q0.x, q0.y, q0.z);
    v = mk3<T>(q0.w, q1.x, q1.y);
    C.m[0][0] = q1.z; C.m[0][1] = q1.w; C.m[0][2] = q2.x;
    C.m[1][0] = q2.y; C.m[1][1] = q2.z; C.m[1][2] = q2.w;
    C.m[2][0] = f.a3[p]; C.m[2][1] = f.a4[p]; C.m[2][2] = f.a5[p];
}
template <class T> PLB_HD void store_xvC(const FramePtr<T>& f, int p, V3<T> x, V3<T> v, const M3<T>& C) {
    f.A0[p] = mk4<T>(x.x, x.y, x.z, v.x);
    f.A1[p] = mk4<T>(v.y, v.z, C.m[0][0], C.m[0][1]);
    f.A2[p] = mk4<T>(C.m[0][2], C.m[1][0], C.m[1][1], C.m[1][2]);
    f.a3[p] = C.m[2][0]; f.a4[p] = C.m[2][1]; f.a5[p] = C.m[2][2];
}
template <class T> PLB_HD M3<T> load_F(const FramePtr<T>& f, int p) {
    Vec4<T> q0 = f.B0[p], q1 = f.B1[p];
    M3<T> F;
    F.m[0][0] = q0.x; F.m[0][1] = q0.y; F.m[0][2] = q0.z; F.m[1][0] = q0.w;
    F.m[1][1] = q1.x; F.m[1][2] = q1.y; F.m[2][0] = q1.z; F.m[2][1] = q1.w;
    F.m[2][2] = f.b2[p];
    return F;
}
template <class T> PLB_HD void store_F(const FramePtr<T>& f, int p, const M3<T>& F) {
    f.B0[p] = mk4<T>(F.m[0][0], F.m[0][1], F.m[0][2], F.m[1][0]);
    f.B1[p] = mk4<T>(F.m[1][1], F.m[1][2], F.m[2][0], F.m[2][1]);
    f.b2[p] = F.m[2][2];
}
template <class T> PLB_HD V3<T> load_x(const FramePtr<T>& f, int p) {
    Vec4<T> q0 = f.A0[p];
    return mk3<T>(q0.x, q0.y, q0.z);
}

// ---- scatter primitives: red.global on the device, plain adds in the sequential host emulation
#if defined(__CUDA_ARCH__)
PLB_D void scatter_add4(Vec4<float>* addr, Vec4<float> v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
PLB_D void scatter_add4(Vec4<double>* addr, Vec4<double> v) {
    double* a = reinterpret_cast<double*>(addr);
    atomicAdd(a + 0, v.x); atomicAdd(a + 1, v.y); atomicAdd(a + 2, v.z); atomicAdd(a + 3, v.w);
}
PLB_D void scatter_add1(float* addr, float v) { atomicAdd(addr, v); }
PLB_D void scatter_add1(double* addr, double v) { atomicAdd(addr, v); }
#elif defined(PLB_WARP_EMUL)
inline std::mutex& emul_scatter_mutex() { static std::mutex m; return m; }
template <class T> inline void scatter_add4(Vec4<T>* addr, Vec4<T> v) {
    std::lock_guard<std::mutex> lock(emul_scatter_mutex());
    addr->x += v.x; addr->y += v.y; addr->z += v.z; addr->w += v.w;
}
template <class T> inline void scatter_add1(T* addr, T v) { std::lock_guard<std::mutex> lock(emul_scatter_mutex()); *addr += v; }
#else
template <class T> inline void scatter_add4(Vec4<T>* addr, Vec4<T> v) { addr->x += v.x; addr->y += v.y; addr->z += v.z; addr->w += v.w; }
template <class T> inline void scatter_add1(T* addr, T v) { *addr += v; }
#endif

#ifdef PLB_INDEX32
PLB_HD long long node_index(int n, int i, int j, int k) { return (long long)((i * n + j) * n + k); }      // n <= 1024: fits 31 bits
#else
PLB_HD long long node_index(int n, int i, int j, int k) { return ((long long)i * n + j) * n + k; }
#endif

// 27-node stencil window of a Vec4 grid: at(i, j, k) = node (b0 + i, b1 + j, b2 + k) of the particle's stencil.  `p` points at
// node (b0, b1, b2) either in the dense global grid (strides n^2, n) or in a shared-memory tile of 8^3 nodes (strides 64, 8)
// that TMA loaded for the CTA (plb_tile.cuh): a generic pointer, so the gather loops have one code path.
template <class T> struct GridView {
    const Vec4<T>* p; int si, sj;
    PLB_HD Vec4<T> at(int i, int j, int k) const { return p[i * si + j * sj + k]; }
};
template <class T> PLB_HD GridView<T> global_view(const Vec4<T>* grid, int n, const int b[3]) {
    GridView<T> v;
    v.p = grid + node_index(n, b[0], b[1], b[2]); v.si = n * n; v.sj = n;
    return v;
}

// Scatter policy used by the two scattering bodies.  Direct: one (vector) atomic per node and particle.
// The CUDA kernels can substitute WarpTileScatter (plb_kernels.cuh), which pre-reduces a warp's contributions.
template <class T> struct DirectScatter {
    Vec4<T>* grid;
    int n;
    PLB_HD void add(int slot, int i, int j, int k, Vec4<T> v) const { (void)slot; scatter_add4(grid + node_index(n, i, j, k), v); }
    PLB_HD void end_plane(int) const {}
};

template <class T> PLB_HD void load_material(const SimConst<T>& P, const Material<T>& mat, int p, T& mu, T& lam, T& ys) {
    mu = mat.mu ? mat.mu[p] : P.mu;
    lam = mat.lam ? mat.lam[p] : P.lam;
    ys = mat.ys ? mat.ys[p] : P.yield_stress;
}

// ================================================================================================
// forward substep
// ================================================================================================
// P2G: F_tmp, SVD, return mapping, stress, 27-node scatter.  `out` may alias nothing (F[f+1] store skipped if !store_F_out).
// register-level core: particle state in, new_F out, 27 contributions through the scatter policy
template <class T, class Sc>
PLB_HD void p2g_core(const SimConst<T>& P, V3<T> x, V3<T> v, const M3<T>& C, const M3<T>& F, T mu, T lam, T ys, M3<T>& new_F, const Sc& sc,
                     SvdRec<T>* svd_out = nullptr) {
    M3<T> affine;
    p2g_particle<T>(P, C, F, mu, lam, ys, new_F, affine, nullptr, svd_out);
    Stencil<T> st = make_stencil(x, P.inv_dx);
    // momentum_o = w_o (p_mass v + affine ((o - fx) dx)) = w_o (m0 + i c0 + j c1 + k c2): the affine part is evaluated
    // incrementally along the three stencil axes (81 + 27 + 9 FMAs instead of 27 mat-vecs)
    V3<T> c0 = mk3<T>(affine.m[0][0] * P.dx, affine.m[1][0] * P.dx, affine.m[2][0] * P.dx);
    V3<T> c1 = mk3<T>(affine.m[0][1] * P.dx, affine.m[1][1] * P.dx, affine.m[2][1] * P.dx);
    V3<T> c2 = mk3<T>(affine.m[0][2] * P.dx, affine.m[1][2] * P.dx, affine.m[2][2] * P.dx);
    V3<T> m0 = P.p_mass * v - (st.fx.x * c0 + st.fx.y * c1 + st.fx.z * c2);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        V3<T> mi = m0 + T(i) * c0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            V3<T> mij = mi + T(j) * c1;
            T wij = st.w[i][0] * st.w[j][1];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                T w = wij * st.w[k][2];
                V3<T> mom = w * (mij + T(k) * c2);
                sc.add((i * 3 + j) * 3 + k, st.b[0] + i, st.b[1] + j, st.b[2] + k, mk4<T>(mom.x, mom.y, mom.z, w * P.p_mass));
            }
        }
        sc.end_plane(i);
    }
}
// svd_keep (optional): planes of the SVD store of this frame; written when the frame's F' is (store_F_out)
template <class T, class Sc>
PLB_HD void p2g_body(int p, const SimConst<T>& P, const FramePtr<T>& in, const FramePtr<T>& out, bool store_F_out,
                     const Material<T>& mat, const Sc& sc, const SvdPtr<T>* svd_keep = nullptr) {
    V3<T> x, v; M3<T> C;
    load_xvC(in, p, x, v, C);
    M3<T> F = load_F(in, p);
    T mu, lam, ys;
    load_material(P, mat, p, mu, lam, ys);
    M3<T> new_F;
    SvdRec<T> rec;
    p2g_core<T, Sc>(P, x, v, C, F, mu, lam, ys, new_F, sc, svd_keep ? &rec : nullptr);
    if (store_F_out) {
        store_F(out, p, new_F);
        if (svd_keep) store_svd(*svd_keep, p, rec);
    }
}
template <class T>
PLB_HD void p2g_body(int p, const SimConst<T>& P, const FramePtr<T>& in, const FramePtr<T>& out, bool store_F_out,
                     const Material<T>& mat, Vec4<T>* grid_in, const SvdPtr<T>* svd_keep = nullptr) {
    DirectScatter<T> sc{grid_in, P.n_grid};
    p2g_body<T, DirectScatter<T>>(p, P, in, out, store_F_out, mat, sc, svd_keep);
}

// grid operator: grid_in -> grid_out; optionally zeroes grid_in for the next scatter
template <class T>
PLB_HD void grid_fwd_body(long long node, const SimConst<T>& P, const PrimSet<T>& prims, const Pose<T>* s0, const Pose<T>* s1,
                          Vec4<T>* grid_in, Vec4<T>* grid_out, bool clear_in) {
    Vec4<T> in4 = grid_in[node];
    const int n = P.n_grid;
    const unsigned un = (unsigned)node, nn = (unsigned)n, row = un / nn;        // n_grid <= 1024: node < 2^30, 32-bit divisions
    int k = (int)(un - row * nn), i = (int)(row / nn), j = (int)(row - (unsigned)i * nn);
    V3<T> v = grid_node_forward<T>(P, prims, s0, s1, i, j, k, in4);
    grid_out[node] = mk4<T>(v.x, v.y, v.z, T(0));
    if (clear_in && (in4.x != T(0) || in4.y != T(0) || in4.z != T(0) || in4.w != T(0)))
        grid_in[node] = mk4<T>(T(0), T(0), T(0), T(0));
}

// G2P: 27-node gather, APIC C, advection
template <class T>
PLB_HD void g2p_core(const SimConst<T>& P, V3<T> x, const Stencil<T>& st, const GridView<T>& gv, V3<T>& nx, V3<T>& nv_out, M3<T>& nC_out) {
    // v' = sum w g;  C' = 4 inv_dx sum w g (x) (o - fx) = 4 inv_dx (sum w g (x) o - v' (x) fx): accumulate sum w g and the
    // three offset-weighted sums (offsets are 0/1/2, so these are adds), one rank-1 correction at the end
    // (written as multiply-add chains: t0 = sum_k w_k g_k and tk = sum_k k w_k g_k per (i, j) column, then four accumulations
    //  with the column weight -- 23 FMA-class instructions per column instead of the ~40 of the product-then-add form)
    V3<T> nv = zero3<T>(), si = zero3<T>(), sj = zero3<T>(), sk = zero3<T>();
    const T wk0 = st.w[0][2], wk1 = st.w[1][2], wk2 = st.w[2][2], wk2x2 = T(2) * st.w[2][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const T wij = st.w[i][0] * st.w[j][1];
            const Vec4<T> a4 = gv.at(i, j, 0), b4 = gv.at(i, j, 1), c4v = gv.at(i, j, 2);
            const V3<T> g0 = mk3<T>(a4.x, a4.y, a4.z), g1 = mk3<T>(b4.x, b4.y, b4.z), g2 = mk3<T>(c4v.x, c4v.y, c4v.z);
            const V3<T> t0 = fma3(wk2, g2, fma3(wk1, g1, wk0 * g0));
            const V3<T> tk = fma3(wk2x2, g2, wk1 * g1);
            nv = fma3(wij, t0, nv);
            sk = fma3(wij, tk, sk);
            if (i > 0) si = fma3(T(i) * wij, t0, si);
            if (j > 0) sj = fma3(T(j) * wij, t0, sj);
        }
    M3<T> nC;
    const T c4 = T(4) * P.inv_dx;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        nC.m[r][0] = c4 * (si[r] - nv[r] * st.fx.x);
        nC.m[r][1] = c4 * (sj[r] - nv[r] * st.fx.y);
        nC.m[r][2] = c4 * (sk[r] - nv[r] * st.fx.z);
    }
    nx = advect(P, x, nv);
    nv_out = nv;
    nC_out = nC;
}
template <class T>
PLB_HD void g2p_core(const SimConst<T>& P, V3<T> x, const Vec4<T>* grid_out, V3<T>& nx, V3<T>& nv_out, M3<T>& nC_out) {
    Stencil<T> st = make_stencil(x, P.inv_dx);
    g2p_core<T>(P, x, st, global_view(grid_out, P.n_grid, st.b), nx, nv_out, nC_out);
}
template <class T>
PLB_HD void g2p_body(int p, const SimConst<T>& P, const FramePtr<T>& in, const FramePtr<T>& out, const Vec4<T>* grid_out) {
    V3<T> nx, nv; M3<T> nC;
    g2p_core<T>(P, load_x(in, p), grid_out, nx, nv, nC);
    store_xvC(out, p, nx, nv, nC);
}

// ================================================================================================
// backward substep (after P2G + grid_fwd were recomputed for frame f)
// ================================================================================================
// g2p.grad: reads adjoint of (x,v,C)[f+1], scatters the adjoint of grid_out, writes the partial x-adjoint of frame f
// into adj_cur.A0 (xyz lanes; the w lane is finished by p2g_bwd_body).
// kStoredNext: (xn, nv) = the position / velocity G2P stored for the next frame.  nv IS the gather sum (g2p stores it
// unchanged) and the clamp masks of the advection can be read off the stored position (0 < x' < x_hi <=> both clamps pass),
// so the first 27-node gather is skipped.  Otherwise (no forward result at hand) the sum is recomputed from grid_out.
template <class T, class Sc, bool kStoredNext = false>
PLB_HD V3<T> g2p_bwd_core(const SimConst<T>& P, V3<T> x, V3<T> gxn, V3<T> gvn, const M3<T>& gCn, const Vec4<T>* grid_out, const Sc& sc,
                          V3<T> xn = V3<T>(), V3<T> nv_stored = V3<T>()) {
    Stencil<T> st = make_stencil(x, P.inv_dx);
    V3<T> nv, gy;
    if (kStoredNext) {
        nv = nv_stored;
        gy = advect_backward_stored(P, xn, gxn);
    } else {
        // recompute new_v = sum w g (clamp masks of the advection; it is also the sum the dpos adjoint needs)
        nv = zero3<T>();
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                V3<T> t0 = zero3<T>();
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    Vec4<T> g4 = grid_out[node_index(P.n_grid, st.b[0] + i, st.b[1] + j, st.b[2] + k)];
                    t0 += st.w[k][2] * mk3<T>(g4.x, g4.y, g4.z);
                }
                nv += (st.w[i][0] * st.w[j][1]) * t0;
            }
        gy = advect_backward(P, x, nv, gxn);
    }
    V3<T> gx = gy;
    V3<T> gv = gvn + P.dt * gy;
    T gw[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int d = 0; d < 3; d++) gw[a][d] = T(0);
    V3<T> gfx = zero3<T>();
    const T c4 = T(4) * P.inv_dx;
    // adjoint of grid_out[o] = w_o h_o with h_o = gv + c4 gC' (o - fx): h is evaluated incrementally along the axes;
    // the adjoint of the 3-D weight is g_v[o] . h_o and the adjoint of dpos sums to c4 gC'^T (sum_o w_o g_v[o])
    V3<T> hc0 = mk3<T>(c4 * gCn.m[0][0], c4 * gCn.m[1][0], c4 * gCn.m[2][0]);
    V3<T> hc1 = mk3<T>(c4 * gCn.m[0][1], c4 * gCn.m[1][1], c4 * gCn.m[2][1]);
    V3<T> hc2 = mk3<T>(c4 * gCn.m[0][2], c4 * gCn.m[1][2], c4 * gCn.m[2][2]);
    V3<T> h0 = gv - (st.fx.x * hc0 + st.fx.y * hc1 + st.fx.z * hc2);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        V3<T> hi = h0 + T(i) * hc0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            V3<T> hij = hi + T(j) * hc1;
            T wij = st.w[i][0] * st.w[j][1];
            T tij = T(0);                                   // sum_k g(weight_ijk) w_k
#pragma unroll
            for (int k = 0; k < 3; k++) {
                T w = wij * st.w[k][2];
                V3<T> h = hij + T(k) * hc2;
                Vec4<T> g4 = grid_out[node_index(P.n_grid, st.b[0] + i, st.b[1] + j, st.b[2] + k)];
                sc.add((i * 3 + j) * 3 + k, st.b[0] + i, st.b[1] + j, st.b[2] + k, mk4<T>(w * h.x, w * h.y, w * h.z, T(0)));
                T gwt = g4.x * h.x + g4.y * h.y + g4.z * h.z;
                tij += gwt * st.w[k][2];
                gw[k][2] += gwt * wij;
            }
            gw[i][0] += tij * st.w[j][1];
            gw[j][1] += tij * st.w[i][0];
        }
        sc.end_plane(i);
    }
    gfx = (-c4) * mTv(gCn, nv);
    gx += stencil_backward(st, gw, gfx, P.inv_dx);
    return gx;
}
// ---- g2p.grad split in two passes (plb_tile.cuh): the gather pass reads grid_out (through a view, possibly a shared-memory
// tile) and produces the partial x-adjoint; the scatter pass only needs the stencil and the coefficients of h_o, so that the
// shared-memory tile can be dead -- and its space reused by the scatter tiles -- before the first contribution is parked.
template <class T> struct G2PBwdCarry { V3<T> h0, hc0, hc1, hc2; };
template <class T, bool kStoredNext>
PLB_HD V3<T> g2p_bwd_gather(const SimConst<T>& P, V3<T> x, const Stencil<T>& st, const GridView<T>& gv, V3<T> gxn, V3<T> gvn, const M3<T>& gCn,
                            V3<T> xn, V3<T> nv_stored, G2PBwdCarry<T>& carry) {
    V3<T> nv, gy;
    if (kStoredNext) {
        nv = nv_stored;
        gy = advect_backward_stored(P, xn, gxn);
    } else {
        nv = zero3<T>();
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                V3<T> t0 = zero3<T>();
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    Vec4<T> g4 = gv.at(i, j, k);
                    t0 += st.w[k][2] * mk3<T>(g4.x, g4.y, g4.z);
                }
                nv += (st.w[i][0] * st.w[j][1]) * t0;
            }
        gy = advect_backward(P, x, nv, gxn);
    }
    V3<T> gx = gy;
    V3<T> gvv = gvn + P.dt * gy;
    T gw[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int d = 0; d < 3; d++) gw[a][d] = T(0);
    const T c4 = T(4) * P.inv_dx;
    carry.hc0 = mk3<T>(c4 * gCn.m[0][0], c4 * gCn.m[1][0], c4 * gCn.m[2][0]);
    carry.hc1 = mk3<T>(c4 * gCn.m[0][1], c4 * gCn.m[1][1], c4 * gCn.m[2][1]);
    carry.hc2 = mk3<T>(c4 * gCn.m[0][2], c4 * gCn.m[1][2], c4 * gCn.m[2][2]);
    carry.h0 = gvv - (st.fx.x * carry.hc0 + st.fx.y * carry.hc1 + st.fx.z * carry.hc2);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        V3<T> hi = carry.h0 + T(i) * carry.hc0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            V3<T> hij = hi + T(j) * carry.hc1;
            T wij = st.w[i][0] * st.w[j][1];
            T tij = T(0);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                V3<T> h = hij + T(k) * carry.hc2;
                Vec4<T> g4 = gv.at(i, j, k);
                T gwt = g4.x * h.x + g4.y * h.y + g4.z * h.z;
                tij += gwt * st.w[k][2];
                gw[k][2] += gwt * wij;
            }
            gw[i][0] += tij * st.w[j][1];
            gw[j][1] += tij * st.w[i][0];
        }
    }
    V3<T> gfx = (-c4) * mTv(gCn, nv);
    gx += stencil_backward(st, gw, gfx, P.inv_dx);
    return gx;
}
template <class T, class Sc>
PLB_HD void g2p_bwd_scatter(const Stencil<T>& st, const G2PBwdCarry<T>& carry, const Sc& sc) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        V3<T> hi = carry.h0 + T(i) * carry.hc0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            V3<T> hij = hi + T(j) * carry.hc1;
            T wij = st.w[i][0] * st.w[j][1];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                T w = wij * st.w[k][2];
                V3<T> h = hij + T(k) * carry.hc2;
                sc.add((i * 3 + j) * 3 + k, st.b[0] + i, st.b[1] + j, st.b[2] + k, mk4<T>(w * h.x, w * h.y, w * h.z, T(0)));
            }
        }
        sc.end_plane(i);
    }
}
// P2G scatter loop alone (the particle part -- p2g_particle -- already ran): 27 contributions through the scatter policy
template <class T, class Sc>
PLB_HD void p2g_scatter(const SimConst<T>& P, const Stencil<T>& st, V3<T> v, const M3<T>& affine, const Sc& sc) {
    V3<T> c0 = mk3<T>(affine.m[0][0] * P.dx, affine.m[1][0] * P.dx, affine.m[2][0] * P.dx);
    V3<T> c1 = mk3<T>(affine.m[0][1] * P.dx, affine.m[1][1] * P.dx, affine.m[2][1] * P.dx);
    V3<T> c2 = mk3<T>(affine.m[0][2] * P.dx, affine.m[1][2] * P.dx, affine.m[2][2] * P.dx);
    V3<T> m0 = P.p_mass * v - (st.fx.x * c0 + st.fx.y * c1 + st.fx.z * c2);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        V3<T> mi = m0 + T(i) * c0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            V3<T> mij = mi + T(j) * c1;
            T wij = st.w[i][0] * st.w[j][1];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                T w = wij * st.w[k][2];
                V3<T> mom = w * (mij + T(k) * c2);
                sc.add((i * 3 + j) * 3 + k, st.b[0] + i, st.b[1] + j, st.b[2] + k, mk4<T>(mom.x, mom.y, mom.z, w * P.p_mass));
            }
        }
        sc.end_plane(i);
    }
}

// fnext: the frame G2P produced from `in` (null: not available, recompute the gather sum)
template <class T, class Sc>
PLB_HD void g2p_bwd_body(int p, const SimConst<T>& P, const FramePtr<T>& in, const FramePtr<T>& adj_next,
                         const FramePtr<T>& adj_cur, const Vec4<T>* grid_out, const Sc& sc, const FramePtr<T>* fnext = nullptr) {
    V3<T> gxn, gvn; M3<T> gCn;
    load_xvC(adj_next, p, gxn, gvn, gCn);
    V3<T> gx;
    if (fnext) {
        Vec4<T> q0 = fnext->A0[p], q1 = fnext->A1[p];
        gx = g2p_bwd_core<T, Sc, true>(P, load_x(in, p), gxn, gvn, gCn, grid_out, sc, mk3<T>(q0.x, q0.y, q0.z), mk3<T>(q0.w, q1.x, q1.y));
    } else {
        gx = g2p_bwd_core<T, Sc, false>(P, load_x(in, p), gxn, gvn, gCn, grid_out, sc);
    }
    adj_cur.A0[p] = mk4<T>(gx.x, gx.y, gx.z, T(0));
}
template <class T>
PLB_HD void g2p_bwd_body(int p, const SimConst<T>& P, const FramePtr<T>& in, const FramePtr<T>& adj_next,
                         const FramePtr<T>& adj_cur, const Vec4<T>* grid_out, Vec4<T>* g_out, const FramePtr<T>* fnext = nullptr) {
    DirectScatter<T> sc{g_out, P.n_grid};
    g2p_bwd_body<T, DirectScatter<T>>(p, P, in, adj_next, adj_cur, grid_out, sc, fnext);
}

// grid_op.grad for one node.  Reads grid_in (forward values) and g_out (adjoint of grid_out); writes g_in (adjoint of
// grid_in).  Pose adjoints are returned through g0/g1/touched for the caller to reduce.
template <class T>
PLB_HD void grid_bwd_body(long long node, const SimConst<T>& P, const PrimSet<T>& prims, const Pose<T>* s0, const Pose<T>* s1,
                          Vec4<T>* grid_in, Vec4<T>* g_out, Vec4<T>* g_in, bool clear,
                          PoseGrad<T>* g0, PoseGrad<T>* g1, unsigned& touched) {
    Vec4<T> in4 = grid_in[node];
    Vec4<T> go = g_out[node];
    const int n = P.n_grid;
    const unsigned un = (unsigned)node, nn = (unsigned)n, row = un / nn;        // n_grid <= 1024: node < 2^30, 32-bit divisions
    int k = (int)(un - row * nn), i = (int)(row / nn), j = (int)(row - (unsigned)i * nn);
    Vec4<T> gi = grid_node_backward<T>(P, prims, s0, s1, i, j, k, in4, mk3<T>(go.x, go.y, go.z), g0, g1, touched);
    g_in[node] = gi;
    if (clear) {
        if (in4.x != T(0) || in4.y != T(0) || in4.z != T(0) || in4.w != T(0)) grid_in[node] = mk4<T>(T(0), T(0), T(0), T(0));
        if (go.x != T(0) || go.y != T(0) || go.z != T(0)) g_out[node] = mk4<T>(T(0), T(0), T(0), T(0));
    }
}

// p2g.grad + svd_grad + compute_F_tmp.grad: gathers g_in at 27 nodes, finishes the adjoint of frame f in adj_cur.
// register-level core: state of frame f, adjoint of F[f+1], partial x-adjoint -> full adjoint of frame f
// kSvdGiven: `svd` holds the forward's decomposition of F_tmp (SVD store) and the Jacobi iteration is not re-run
// kTwoPhase (with kSvdGiven): the forward particle math is cheap once the SVD is given (~250 instructions), so it is run
// twice -- before the 27-node gather only for `affine`, and again after it for the adjoint -- instead of keeping its ~58
// intermediate values (P2GState) in registers across the gather loop, which is where the backward kernels' register peak is.
template <class T, bool kSvdGiven = false, bool kTwoPhase = false>
PLB_HD void p2g_bwd_core(const SimConst<T>& P, const Stencil<T>& st, const GridView<T>& gin_view, V3<T> v, const M3<T>& C, const M3<T>& F, T mu, T lam, T ys,
                         const M3<T>& gF_next, V3<T> gx_partial, V3<T>& gx_out, V3<T>& gv_out, M3<T>& gC, M3<T>& gF,
                         SvdRec<T>* svd = nullptr, const FramePtr<T>* gF_late = nullptr, int p_late = 0, const FramePtr<T>* state_late = nullptr) {
    // (kTwoPhase: after the gather loop the adjoint of F[f+1] is loaded from gF_late instead of being passed in gF_next, and
    //  C, F are loaded again from state_late -- L1 hits -- so that neither occupies registers across the loop)
    M3<T> new_F, affine;
    P2GState<T> keep;
    if (kSvdGiven && kTwoPhase) p2g_particle<T, true>(P, C, F, mu, lam, ys, new_F, affine, nullptr, svd);
    else p2g_particle<T, kSvdGiven>(P, C, F, mu, lam, ys, new_F, affine, &keep, svd);
    // forward node value: w_o (m_o, p_mass) with m_o = p_mass v + affine (o - fx) dx, evaluated incrementally.
    // With a_o = adjoint of the node momentum: g(weight_o) = a_o . m_o + b_o p_mass;  g(v) = p_mass sum w a;
    // g(affine) = (sum w a (x) o - (sum w a) (x) fx) dx;  g(fx) -= dx affine^T (sum w a)  -- the last two leave the loop.
    V3<T> c0 = mk3<T>(affine.m[0][0] * P.dx, affine.m[1][0] * P.dx, affine.m[2][0] * P.dx);
    V3<T> c1 = mk3<T>(affine.m[0][1] * P.dx, affine.m[1][1] * P.dx, affine.m[2][1] * P.dx);
    V3<T> c2 = mk3<T>(affine.m[0][2] * P.dx, affine.m[1][2] * P.dx, affine.m[2][2] * P.dx);
    V3<T> m0 = P.p_mass * v - (st.fx.x * c0 + st.fx.y * c1 + st.fx.z * c2);
    V3<T> gv = zero3<T>();                 // sum w a
    V3<T> si = zero3<T>(), sj = zero3<T>(), sk = zero3<T>();     // sum w a * {i, j, k}
    T gw[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int d = 0; d < 3; d++) gw[a][d] = T(0);
    // per (i, j) column, multiply-add chains over k (as in g2p_core): A0 = sum_k w_k a_k, Ak = sum_k k w_k a_k feed the four
    // accumulators with the column weight; the weight adjoints take gwt_k = a_k . m_ijk + b_k p_mass
    const T wk0 = st.w[0][2], wk1 = st.w[1][2], wk2 = st.w[2][2], wk2x2 = T(2) * st.w[2][2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        V3<T> mi = m0 + T(i) * c0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const V3<T> mij0 = mi + T(j) * c1, mij1 = mij0 + c2, mij2 = mij1 + c2;
            const T wij = st.w[i][0] * st.w[j][1];
            const Vec4<T> q0 = gin_view.at(i, j, 0), q1 = gin_view.at(i, j, 1), q2 = gin_view.at(i, j, 2);
            const V3<T> a0 = mk3<T>(q0.x, q0.y, q0.z), a1 = mk3<T>(q1.x, q1.y, q1.z), a2 = mk3<T>(q2.x, q2.y, q2.z);
            const V3<T> A0 = fma3(wk2, a2, fma3(wk1, a1, wk0 * a0));
            const V3<T> Ak = fma3(wk2x2, a2, wk1 * a1);
            gv = fma3(wij, A0, gv);
            sk = fma3(wij, Ak, sk);
            if (i > 0) si = fma3(T(i) * wij, A0, si);
            if (j > 0) sj = fma3(T(j) * wij, A0, sj);
            const T gw0 = dot(a0, mij0) + q0.w * P.p_mass, gw1 = dot(a1, mij1) + q1.w * P.p_mass, gw2 = dot(a2, mij2) + q2.w * P.p_mass;
            const T tij = gw0 * wk0 + gw1 * wk1 + gw2 * wk2;
            gw[0][2] += gw0 * wij; gw[1][2] += gw1 * wij; gw[2][2] += gw2 * wij;
            gw[i][0] += tij * st.w[j][1];
            gw[j][1] += tij * st.w[i][0];
        }
    }
    M3<T> g_aff;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        g_aff.m[r][0] = (si[r] - gv[r] * st.fx.x) * P.dx;
        g_aff.m[r][1] = (sj[r] - gv[r] * st.fx.y) * P.dx;
        g_aff.m[r][2] = (sk[r] - gv[r] * st.fx.z) * P.dx;
    }
    V3<T> gfx = (-P.dx) * mTv(affine, gv);
    gv = P.p_mass * gv;
    gx_out = stencil_backward(st, gw, gfx, P.inv_dx) + gx_partial;
    gv_out = gv;
    if (kSvdGiven && kTwoPhase) {
        // second run of the forward particle math, on values the compiler cannot tie to the first run
        SvdRec<T> r2 = *svd;
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) { opaque(r2.U.m[i][j]); opaque(r2.V.m[i][j]); }
            opaque(r2.sig[i]);
        }
        M3<T> nf2, aff2;
        V3<T> x2, v2; M3<T> C2;
        load_xvC(*state_late, p_late, x2, v2, C2);
        M3<T> F2 = load_F(*state_late, p_late);
        p2g_particle<T, true>(P, C2, F2, mu, lam, ys, nf2, aff2, &keep, &r2);
        p2g_particle_backward<T>(P, C2, F2, mu, lam, keep, g_aff, load_F(*gF_late, p_late), gC, gF);
        return;
    }
    p2g_particle_backward<T>(P, C, F, mu, lam, keep, g_aff, gF_next, gC, gF);
}
template <class T, bool kSvdGiven = false, bool kTwoPhase = false>
PLB_HD void p2g_bwd_core(const SimConst<T>& P, V3<T> x, V3<T> v, const M3<T>& C, const M3<T>& F, T mu, T lam, T ys, const Vec4<T>* g_in,
                         const M3<T>& gF_next, V3<T> gx_partial, V3<T>& gx_out, V3<T>& gv_out, M3<T>& gC, M3<T>& gF,
                         SvdRec<T>* svd = nullptr, const FramePtr<T>* gF_late = nullptr, int p_late = 0, const FramePtr<T>* state_late = nullptr) {
    Stencil<T> st = make_stencil(x, P.inv_dx);
    p2g_bwd_core<T, kSvdGiven, kTwoPhase>(P, st, global_view(g_in, P.n_grid, st.b), v, C, F, mu, lam, ys, gF_next, gx_partial, gx_out, gv_out, gC, gF,
                                          svd, gF_late, p_late, state_late);
}
template <class T, bool kSvdGiven = false>
PLB_HD void p2g_bwd_body(int p, const SimConst<T>& P, const FramePtr<T>& in, const FramePtr<T>& adj_next,
                         const FramePtr<T>& adj_cur, const Material<T>& mat, const Vec4<T>* g_in, const SvdPtr<T>* svd_kept = nullptr) {
    V3<T> x, v; M3<T> C;
    load_xvC(in, p, x, v, C);
    M3<T> F = load_F(in, p);
    T mu, lam, ys;
    load_material(P, mat, p, mu, lam, ys);
    Vec4<T> part = adj_cur.A0[p];                           // partial x-adjoint from g2p_bwd_body
    V3<T> gx, gv; M3<T> gC, gF;
    SvdRec<T> rec;
    if (kSvdGiven) rec = load_svd(*svd_kept, p);
    p2g_bwd_core<T, kSvdGiven>(P, x, v, C, F, mu, lam, ys, g_in, load_F(adj_next, p), mk3<T>(part.x, part.y, part.z), gx, gv, gC, gF, &rec);
    store_xvC(adj_cur, p, gx, gv, gC);
    store_F(adj_cur, p, gF);
}

// ================================================================================================
// loss (plb/engine/losses/loss.py:116-153,186-237 ; mpm_simulator.py:382-392)
// ================================================================================================
template <class T>
PLB_HD void loss_mass_body(int p, const SimConst<T>& P, const FramePtr<T>& in, T* grid_mass) {
    V3<T> x = load_x(in, p);
    Stencil<T> st = make_stencil(x, P.inv_dx);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                T w = st.w[i][0] * st.w[j][1] * st.w[k][2];
                scatter_add1(grid_mass + node_index(P.n_grid, st.b[0] + i, st.b[1] + j, st.b[2] + k), w * P.p_mass);
            }
}

struct LossWeights { double sdf, density, contact; int soft; };

// Adjoint of the per-step loss wrt x[f] (added into adj.A0 xyz) and the primitive poses of frame f.
// min_dist[k] (forward value of primitive k's min distance) seeds the contact term; with the reference's
// atomic-min-as-add autodiff every particle with sdf > 0 receives 2 * min_dist * w_contact (contact_all = 1),
// contact_all = 0 restricts it to the arg-min particle(s) (d == min_dist).
template <class T>
PLB_HD void loss_bwd_body(int p, const SimConst<T>& P, const FramePtr<T>& in, const FramePtr<T>& adj,
                          const T* grid_mass, const T* target, const T* target_sdf, T w_sdf, T w_density, T w_contact,
                          const PrimSet<T>& prims, const Pose<T>* s0, const double* min_dist, int contact_all,
                          PoseGrad<T>* gpose, unsigned& touched, int soft = 0, const double* soft_norm = nullptr) {
    V3<T> x = load_x(in, p);
    Stencil<T> st = make_stencil(x, P.inv_dx);
    T gw[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int d = 0; d < 3; d++) gw[a][d] = T(0);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                long long node = node_index(P.n_grid, st.b[0] + i, st.b[1] + j, st.b[2] + k);
                T diff = grid_mass[node] - target[node];
                T gm = w_density * sgn0(diff) + w_sdf * target_sdf[node];
                T gwt = gm * P.p_mass;
                gw[i][0] += gwt * st.w[j][1] * st.w[k][2];
                gw[j][1] += gwt * st.w[i][0] * st.w[k][2];
                gw[k][2] += gwt * st.w[i][0] * st.w[j][1];
            }
    V3<T> gx = stencil_backward(st, gw, zero3<T>(), P.inv_dx);
    touched = 0;
    for (int k = 0; k < P.n_prim; k++) {
        if (!prims.s[k].movable) continue;
        T d = prim_sdf(prims.s[k], s0[k], x);
        if (!(T(0) < d)) continue;                           // max(sdf, 0): gradient only where 0 < sdf
        T g;
        if (soft) {
            // min_dist = S1 / S0, S1 = sum d sw, S0 = sum sw (loss.py:116-135).  min_dist[k] holds S1, soft_norm[k] holds S0.
            // d(min_dist^2 w_c)/dd_i = g_md [ (sw + d sw') / S0  -  (S1 / S0^2) sw' ],  sw' = -2e4 d sw^2
            double S0 = soft_norm[k], md_ = min_dist[k] / S0, dd = (double)d;
            double sw = 1.0 / (1.0 + dd * dd * 10000.0), dsw = -20000.0 * dd * sw * sw;
            double g_md = 2.0 * md_ * (double)w_contact;
            g = (T)(g_md * ((sw + dd * dsw) / S0 - md_ / S0 * dsw));
        } else {
            T md = (T)min_dist[k];
            if (!contact_all && d > md) continue;
            g = T(2) * md * w_contact;
        }
        if (g != T(0)) {
            prim_sdf_vjp(prims.s[k], s0[k], x, g, gpose[k], &gx);
            touched |= 1u << k;
        }
    }
    Vec4<T> a0 = adj.A0[p];
    adj.A0[p] = mk4<T>(a0.x + gx.x, a0.y + gx.y, a0.z + gx.z, a0.w);
}

}  // namespace plb
