// Warp-level scatter kernels (plasticinelab_b200/csrc/plb_warp.cuh) run on the CPU through tests/host/warp_emul.hpp and
// compared with the direct per-particle bodies -- TEST CODE ONLY (never linked into the product library).
//
// wemul_check runs two forward substeps and their adjoints twice: once with the sequential direct-scatter bodies (the
// path tests/test_host_emulation.py already pins against the float64 oracle) and once with the thread-level scatter
// kernels t_p2g / t_g2p_p2g / t_g2p_bwd / t_p2g_bwd_g2p_bwd / t_loss_mass on emulated warps, and reports the largest
// deviation of every output (grids, frames, adjoints), normalised by the largest reference magnitude.
#define PLB_WARP_EMUL 1
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "warp_emul.hpp"
#include "../../plasticinelab_b200/csrc/plb_setup.hpp"
#include "../../plasticinelab_b200/csrc/plb_warp.cuh"

using namespace plb;

namespace {

template <class T> double rel_dev(const T* a, const T* b, size_t n) {
    double d = 0, m = 0;
    for (size_t i = 0; i < n; i++) { d = std::max(d, std::fabs((double)a[i] - (double)b[i])); m = std::max(m, std::fabs((double)b[i])); }
    return m > 0 ? d / m : d;
}
template <class T> double rel_dev(const std::vector<Vec4<T>>& a, const std::vector<Vec4<T>>& b) {
    return rel_dev(reinterpret_cast<const T*>(a.data()), reinterpret_cast<const T*>(b.data()), a.size() * 4);
}

template <class T> struct World {
    plb_config cfg; SimConst<T> P; PrimSet<T> prims; long long n_pad, n_nodes;
    Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    std::vector<T> f[3], adj[3];
    std::vector<Vec4<T>> grid_in, grid_out[2], g_out, g_in;
    std::vector<unsigned char> flags;
    Material<T> mat{nullptr, nullptr, nullptr};
    World(const plb_config& c, const plb_primitive_desc* pd, double softness, const double* pose0, const double* pose1) : cfg(c) {
        P = make_simconst<T>(c);
        for (int k = 0; k < c.n_primitives; k++) { prims.s[k] = make_primstatic<T>(pd[k], softness); s0[k] = load_pose<T>(pose0 + k * 8); s1[k] = load_pose<T>(pose1 + k * 8); }
        n_pad = ((long long)c.n_particles + 31) / 32 * 32;
        n_nodes = (long long)c.n_grid * c.n_grid * c.n_grid;
        for (int i = 0; i < 3; i++) { f[i].assign(24 * n_pad, T(0)); adj[i].assign(24 * n_pad, T(0)); }
        Vec4<T> z = mk4<T>(T(0), T(0), T(0), T(0));
        grid_in.assign(n_nodes, z); grid_out[0].assign(n_nodes, z); grid_out[1].assign(n_nodes, z); g_out.assign(n_nodes, z); g_in.assign(n_nodes, z);
        flags.assign((size_t)(c.n_grid / 4) * (c.n_grid / 4) * (c.n_grid / 4), 0);
    }
    FramePtr<T> fr(int i) { return frame_at(f[i].data(), 0, n_pad); }
    FramePtr<T> ad(int i) { return frame_at(adj[i].data(), 0, n_pad); }
    void pack(std::vector<T>& buf, const double* x, const double* v, const double* F, const double* C) {
        FramePtr<T> q = frame_at(buf.data(), 0, n_pad);
        for (int p = 0; p < cfg.n_particles; p++) {
            M3<T> CC, FF;
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { CC.m[i][j] = (T)C[p * 9 + i * 3 + j]; FF.m[i][j] = (T)F[p * 9 + i * 3 + j]; }
            store_xvC(q, p, mk3<T>((T)x[p * 3], (T)x[p * 3 + 1], (T)x[p * 3 + 2]), mk3<T>((T)v[p * 3], (T)v[p * 3 + 1], (T)v[p * 3 + 2]), CC);
            store_F(q, p, FF);
        }
    }
    void grid_op(std::vector<Vec4<T>>& out) {        // grid operator on grid_in -> out, then clear grid_in
        for (long long n = 0; n < n_nodes; n++) grid_fwd_body<T>(n, P, prims, s0, s1, grid_in.data(), out.data(), true);
    }
    void grid_adj(const std::vector<Vec4<T>>& fwd_in) {   // g_out -> g_in with the forward grid `fwd_in`; clears g_out
        std::vector<Vec4<T>> in = fwd_in;
        for (long long n = 0; n < n_nodes; n++) {
            PoseGrad<T> g0[PLB_MAX_PRIM], g1[PLB_MAX_PRIM];
            for (int k = 0; k < cfg.n_primitives; k++) { g0[k].clear(); g1[k].clear(); }
            unsigned touched = 0;
            grid_bwd_body<T>(n, P, prims, s0, s1, in.data(), g_out.data(), g_in.data(), true, g0, g1, touched);
        }
    }
    // every warp of the launch, one after the other; fn(p, lane, tile)
    template <class Pay, class Fn> void launch(int tile_elems, Fn fn) {
        const int n_warps = (int)(n_pad / 32);
        for (int w = 0; w < n_warps; w++) {
            std::vector<Pay> tile(tile_elems);
            for (auto& e : tile) std::memset(&e, 0xFF, sizeof(Pay));      // poison (NaN): whatever is read must have been written
            run_warp([&](int lane) { fn(w * 32 + lane, lane, tile.data()); });
        }
    }
};

template <class T, bool kTwoPhase>
int check(const plb_config* c, const plb_primitive_desc* pd, double softness, const double* x, const double* v, const double* F,
          const double* C, const double* pose0, const double* pose1, const double* gx, const double* gv, const double* gF,
          const double* gC, int stored_next, int flush_mode, int svd_store, double* out) {
    World<T> R(*c, pd, softness, pose0, pose1), W(*c, pd, softness, pose0, pose1);
    const int n = c->n_particles;
    const int tile_elems = kTileVec4;
    int o = 0;
    // ------------------------------------------------ reference: sequential direct-scatter bodies
    R.pack(R.f[0], x, v, F, C);
    for (int p = 0; p < n; p++) p2g_body<T>(p, R.P, R.fr(0), R.fr(1), true, R.mat, R.grid_in.data());
    std::vector<Vec4<T>> ref_in0 = R.grid_in;
    R.grid_op(R.grid_out[0]);
    for (int p = 0; p < n; p++) g2p_body<T>(p, R.P, R.fr(0), R.fr(1), R.grid_out[0].data());
    for (int p = 0; p < n; p++) p2g_body<T>(p, R.P, R.fr(1), R.fr(2), true, R.mat, R.grid_in.data());
    std::vector<Vec4<T>> ref_in1 = R.grid_in;
    R.grid_op(R.grid_out[1]);
    for (int p = 0; p < n; p++) g2p_body<T>(p, R.P, R.fr(1), R.fr(2), R.grid_out[1].data());
    // adjoint of frame 2 given; substep 1 then substep 0 (g2p part)
    R.pack(R.adj[2], gx, gv, gF, gC);
    FramePtr<T> rf2 = R.fr(2);
    for (int p = 0; p < n; p++) g2p_bwd_body<T>(p, R.P, R.fr(1), R.ad(2), R.ad(1), R.grid_out[1].data(), R.g_out.data(), stored_next ? &rf2 : nullptr);
    std::vector<Vec4<T>> ref_gout1 = R.g_out;
    double stored_vs_recomputed = 0;
    {   // the successor-frame form of g2p.grad (clamp masks + gather sum from the stored frame) equals the recomputing form
        World<T> Q(*c, pd, softness, pose0, pose1);
        Q.pack(Q.adj[2], gx, gv, gF, gC);
        for (int p = 0; p < n; p++) g2p_bwd_body<T>(p, R.P, R.fr(1), Q.ad(2), Q.ad(1), R.grid_out[1].data(), Q.g_out.data(), stored_next ? nullptr : &rf2);
        stored_vs_recomputed = std::max(rel_dev(Q.g_out, ref_gout1), rel_dev(Q.adj[1].data(), R.adj[1].data(), Q.adj[1].size()));
    }
    R.grid_adj(ref_in1);
    std::vector<T> ref_adj1_partial = R.adj[1];
    for (int p = 0; p < n; p++) p2g_bwd_body<T>(p, R.P, R.fr(1), R.ad(2), R.ad(1), R.mat, R.g_in.data());
    FramePtr<T> rf1 = R.fr(1);
    for (int p = 0; p < n; p++) g2p_bwd_body<T>(p, R.P, R.fr(0), R.ad(1), R.ad(0), R.grid_out[0].data(), R.g_out.data(), &rf1);
    std::vector<Vec4<T>> ref_gout0 = R.g_out;

    // ------------------------------------------------ warp kernels
    W.pack(W.f[0], x, v, F, C);
    // SVD store: two frame slots of 21 scalars per particle (poisoned: whatever the backward loads must have been stored)
    std::vector<T> svd((size_t)2 * kSvdScalars * W.n_pad);
    std::memset(svd.data(), 0xFF, svd.size() * sizeof(T));
    SvdPtr<T> sv0 = svd_at(svd.data(), 0, W.n_pad), sv1 = svd_at(svd.data(), 1, W.n_pad);
    // (1) P2G of substep 0
    W.template launch<Vec4<T>>(tile_elems, [&](int p, int lane, Vec4<T>* tile) {
        t_p2g<T>(p, lane, tile, W.P, W.fr(0), W.fr(1), true, W.mat, W.grid_in.data(), W.flags.data(), flush_mode, svd_store ? &sv0 : nullptr);
    });
    out[o++] = rel_dev(W.grid_in, ref_in0);
    {   // flags: exactly the blocks touched by a particle stencil
        std::vector<unsigned char> fl(W.flags.size(), 0);
        for (int p = 0; p < n; p++) mark_blocks<T>(W.P, load_x(W.fr(0), p), fl.data());
        out[o++] = (fl == W.flags) ? 0.0 : 1.0;
    }
    W.grid_op(W.grid_out[0]);
    // (2) fused G2P(0) + P2G(1)
    W.template launch<Vec4<T>>(tile_elems, [&](int p, int lane, Vec4<T>* tile) {
        t_g2p_p2g<T>(p, lane, tile, W.P, W.fr(0), W.fr(1), W.fr(2), W.mat, W.grid_out[0].data(), W.grid_in.data(), nullptr, flush_mode, svd_store ? &sv1 : nullptr);
    });
    out[o++] = rel_dev(W.grid_in, ref_in1);
    out[o++] = rel_dev(W.f[1].data(), R.f[1].data(), W.f[1].size());
    {   // F planes of frame 2 (x,v,C of frame 2 are produced by the closing G2P below)
        double d = 0;
        for (int p = 0; p < n; p++) { M3<T> a = load_F(W.fr(2), p), b = load_F(R.fr(2), p); for (int i = 0; i < 9; i++) d = std::max(d, std::fabs((double)a.m[i / 3][i % 3] - (double)b.m[i / 3][i % 3])); }
        out[o++] = d;
    }
    W.grid_op(W.grid_out[1]);
    for (int p = 0; p < n; p++) g2p_body<T>(p, W.P, W.fr(1), W.fr(2), W.grid_out[1].data());
    // (3) g2p.grad of substep 1
    W.pack(W.adj[2], gx, gv, gF, gC);
    FramePtr<T> wf2 = W.fr(2);
    W.template launch<Vec4<T>>(tile_elems, [&](int p, int lane, Vec4<T>* tile) {
        t_g2p_bwd<T>(p, lane, tile, W.P, W.fr(1), stored_next ? &wf2 : nullptr, W.ad(2), W.ad(1), W.grid_out[1].data(), W.g_out.data(), flush_mode);
    });
    out[o++] = rel_dev(W.g_out, ref_gout1);
    out[o++] = rel_dev(W.adj[1].data(), ref_adj1_partial.data(), W.adj[1].size());
    W.grid_adj(ref_in1);
    // (4) fused p2g.grad(1) + g2p.grad(0)
    W.template launch<Vec4<T>>(tile_elems, [&](int p, int lane, Vec4<T>* tile) {
        if (svd_store)
            t_p2g_bwd_g2p_bwd<T, true, kTwoPhase>(p, lane, tile, W.P, W.fr(1), W.fr(0), W.ad(2), W.ad(1), W.mat, W.g_in.data(), W.grid_out[0].data(), W.g_out.data(), flush_mode, &sv1);
        else
            t_p2g_bwd_g2p_bwd<T, false>(p, lane, tile, W.P, W.fr(1), W.fr(0), W.ad(2), W.ad(1), W.mat, W.g_in.data(), W.grid_out[0].data(), W.g_out.data(), flush_mode);
    });
    out[o++] = rel_dev(W.g_out, ref_gout0);
    {   // dF[1] (F planes of adj 1) and the partial x-adjoint of frame 0 (written into adj 2's A0 plane)
        double d = 0, m = 0, dx = 0, mx = 0;
        for (int p = 0; p < n; p++) {
            M3<T> a = load_F(W.ad(1), p), b = load_F(R.ad(1), p);
            for (int i = 0; i < 9; i++) { d = std::max(d, std::fabs((double)a.m[i / 3][i % 3] - (double)b.m[i / 3][i % 3])); m = std::max(m, std::fabs((double)b.m[i / 3][i % 3])); }
            Vec4<T> qa = W.ad(2).A0[p], qb = R.ad(0).A0[p];
            dx = std::max({dx, std::fabs((double)qa.x - (double)qb.x), std::fabs((double)qa.y - (double)qb.y), std::fabs((double)qa.z - (double)qb.z)});
            mx = std::max({mx, std::fabs((double)qb.x), std::fabs((double)qb.y), std::fabs((double)qb.z)});
        }
        out[o++] = m > 0 ? d / m : d;
        out[o++] = mx > 0 ? dx / mx : dx;
    }
    // (5) loss mass scatter
    {
        std::vector<T> gm_ref(R.n_nodes, T(0)), gm(W.n_nodes, T(0));
        for (int p = 0; p < n; p++) loss_mass_body<T>(p, R.P, R.fr(0), gm_ref.data());
        W.template launch<T>(kTileVec4, [&](int p, int lane, T* tile) { t_loss_mass<T>(p, lane, tile, W.P, W.fr(0), gm.data()); });
        out[o++] = rel_dev(gm.data(), gm_ref.data(), gm.size());
    }
    out[o++] = stored_vs_recomputed;
    return o;
}

// grid adjoint: t_grid_bwd_node on emulated warps (32 consecutive nodes per warp) vs the array form grid_bwd_body
template <class T>
int check_grid_bwd(const plb_config* c, const plb_primitive_desc* pd, double softness, const double* x, const double* v, const double* F,
                   const double* C, const double* pose0, const double* pose1, const double* gout4, double* out) {
    World<T> R(*c, pd, softness, pose0, pose1), W(*c, pd, softness, pose0, pose1);
    const int n = c->n_particles, np = c->n_primitives;
    for (World<T>* w : {&R, &W}) {
        w->pack(w->f[0], x, v, F, C);
        for (int p = 0; p < n; p++) p2g_body<T>(p, w->P, w->fr(0), w->fr(1), true, w->mat, w->grid_in.data());
        for (long long i = 0; i < w->n_nodes; i++)
            w->g_out[i] = mk4<T>((T)gout4[i * 4], (T)gout4[i * 4 + 1], (T)gout4[i * 4 + 2], T(0));
    }
    std::vector<double> ref(2 * PLB_MAX_PRIM * 8, 0.0), got(2 * PLB_MAX_PRIM * 8, 0.0);
    for (long long node = 0; node < R.n_nodes; node++) {
        PoseGrad<T> g0[PLB_MAX_PRIM], g1[PLB_MAX_PRIM];
        for (int k = 0; k < np; k++) { g0[k].clear(); g1[k].clear(); }
        unsigned touched = 0;
        grid_bwd_body<T>(node, R.P, R.prims, R.s0, R.s1, R.grid_in.data(), R.g_out.data(), R.g_in.data(), true, g0, g1, touched);
        for (int k = 0; k < np; k++) {
            if (!((touched >> k) & 1u)) continue;
            const PoseGrad<T>* gg[2] = {&g0[k], &g1[k]};
            for (int w = 0; w < 2; w++) {
                double* d = ref.data() + ((size_t)w * PLB_MAX_PRIM + k) * 8;
                d[0] += gg[w]->pos.x; d[1] += gg[w]->pos.y; d[2] += gg[w]->pos.z; d[3] += gg[w]->rot.w; d[4] += gg[w]->rot.x;
                d[5] += gg[w]->rot.y; d[6] += gg[w]->rot.z; d[7] += gg[w]->gap;
            }
        }
    }
    // only warps that contain an active node are emulated (the rest return zeros by construction: no mass -> no gradient)
    for (long long w0 = 0; w0 < W.n_nodes; w0 += 32) {
        bool any = false;
        for (int l = 0; l < 32 && w0 + l < W.n_nodes; l++) any = any || W.grid_in[w0 + l].w != T(0) || W.g_out[w0 + l].x != T(0) || W.g_out[w0 + l].y != T(0) || W.g_out[w0 + l].z != T(0);
        if (!any) continue;
        run_warp([&](int lane) {
            const long long node = w0 + lane;
            t_grid_bwd_node<T>(node < W.n_nodes, true, node, lane, W.P, W.prims, W.s0, W.s1, W.grid_in.data(), W.g_out.data(), W.g_in.data(), true,
                               got.data(), 0);
        });
    }
    int o = 0;
    out[o++] = rel_dev(W.g_in, R.g_in);
    out[o++] = rel_dev(got.data(), ref.data(), got.size());
    {   // both cleared grid_in / g_out
        double m = 0;
        for (long long i = 0; i < W.n_nodes; i++) m = std::max({m, std::fabs((double)W.grid_in[i].w), std::fabs((double)W.g_out[i].x), std::fabs((double)W.g_out[i].y), std::fabs((double)W.g_out[i].z)});
        out[o++] = m;
    }
    double nrm = 0;
    for (double r : ref) nrm = std::max(nrm, std::fabs(r));
    out[o++] = nrm;                 // (reported so the test can insist that the contact branch actually ran)
    return o;
}

}  // namespace

extern "C" int wemul_check(int dtype, int two_phase, const plb_config* c, const plb_primitive_desc* pd, double softness, const double* x,
                           const double* v, const double* F, const double* C, const double* pose0, const double* pose1, const double* gx,
                           const double* gv, const double* gF, const double* gC, int stored_next, int flush_mode, int svd_store, double* out) {
    if (dtype == PLB_F32)
        return two_phase ? check<float, true>(c, pd, softness, x, v, F, C, pose0, pose1, gx, gv, gF, gC, stored_next, flush_mode, svd_store, out)
                     : check<float, false>(c, pd, softness, x, v, F, C, pose0, pose1, gx, gv, gF, gC, stored_next, flush_mode, svd_store, out);
    return two_phase ? check<double, true>(c, pd, softness, x, v, F, C, pose0, pose1, gx, gv, gF, gC, stored_next, flush_mode, svd_store, out)
                 : check<double, false>(c, pd, softness, x, v, F, C, pose0, pose1, gx, gv, gF, gC, stored_next, flush_mode, svd_store, out);
}

extern "C" int wemul_check_grid_bwd(int dtype, const plb_config* c, const plb_primitive_desc* pd, double softness, const double* x, const double* v,
                                    const double* F, const double* C, const double* pose0, const double* pose1, const double* gout4, double* out) {
    return dtype == PLB_F32 ? check_grid_bwd<float>(c, pd, softness, x, v, F, C, pose0, pose1, gout4, out)
                            : check_grid_bwd<double>(c, pd, softness, x, v, F, C, pose0, pose1, gout4, out);
}
