// Host emulation of the CUDA kernels' per-thread bodies -- TEST CODE ONLY (never linked into the product library).
//
// The build box has no GPU.  The per-particle / per-node bodies in plasticinelab_b200/csrc/plb_bodies.cuh are
// host+device templates; this file loops them sequentially on the CPU, with the scatter primitives turned into
// plain adds, so the hand-derived adjoints and the packed HBM layout can be checked against the float64 oracle
// here (tests/test_host_emulation.py).  The GPU tests then check the real kernels against the same oracle.
#include <cstring>
#include <vector>
#include "../../plasticinelab_b200/csrc/plb_setup.hpp"

using namespace plb;

namespace {

template <class T> struct Host {
    plb_config cfg;
    SimConst<T> P;
    PrimSet<T> prims;
    long long n_pad, n_nodes;
    std::vector<T> frame_in, frame_out, adj_next, adj_cur;
    std::vector<Vec4<T>> grid_in, grid_out, g_out, g_in;

    Host(const plb_config& c, const plb_primitive_desc* pd, double softness) : cfg(c) {
        P = make_simconst<T>(c);
        for (int k = 0; k < c.n_primitives; k++) prims.s[k] = make_primstatic<T>(pd[k], softness);
        n_pad = ((long long)c.n_particles + 31) / 32 * 32;
        n_nodes = (long long)c.n_grid * c.n_grid * c.n_grid;
        frame_in.assign(24 * n_pad, T(0)); frame_out.assign(24 * n_pad, T(0));
        adj_next.assign(24 * n_pad, T(0)); adj_cur.assign(24 * n_pad, T(0));
        Vec4<T> z = mk4<T>(T(0), T(0), T(0), T(0));
        grid_in.assign(n_nodes, z); grid_out.assign(n_nodes, z); g_out.assign(n_nodes, z); g_in.assign(n_nodes, z);
    }
    void pack(std::vector<T>& fr, const double* x, const double* v, const double* F, const double* C) {
        FramePtr<T> f = frame_at(fr.data(), 0, n_pad);
        for (int p = 0; p < cfg.n_particles; p++) {
            V3<T> xx = mk3<T>((T)x[p * 3], (T)x[p * 3 + 1], (T)x[p * 3 + 2]);
            V3<T> vv = mk3<T>((T)v[p * 3], (T)v[p * 3 + 1], (T)v[p * 3 + 2]);
            M3<T> CC, FF;
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { CC.m[i][j] = (T)C[p * 9 + i * 3 + j]; FF.m[i][j] = (T)F[p * 9 + i * 3 + j]; }
            store_xvC(f, p, xx, vv, CC);
            store_F(f, p, FF);
        }
    }
    void unpack(std::vector<T>& fr, double* x, double* v, double* F, double* C) {
        FramePtr<T> f = frame_at(fr.data(), 0, n_pad);
        for (int p = 0; p < cfg.n_particles; p++) {
            V3<T> xx, vv; M3<T> CC;
            load_xvC(f, p, xx, vv, CC);
            M3<T> FF = load_F(f, p);
            for (int i = 0; i < 3; i++) { x[p * 3 + i] = xx[i]; v[p * 3 + i] = vv[i]; }
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { C[p * 9 + i * 3 + j] = CC.m[i][j]; F[p * 9 + i * 3 + j] = FF.m[i][j]; }
        }
    }
    void poses(const double* p0, const double* p1, Pose<T>* s0, Pose<T>* s1) {
        for (int k = 0; k < cfg.n_primitives; k++) { s0[k] = load_pose<T>(p0 + k * 8); s1[k] = load_pose<T>(p1 + k * 8); }
    }
    void forward_grid(const Pose<T>* s0, const Pose<T>* s1, bool store_F) {
        Material<T> mat{nullptr, nullptr, nullptr};
        FramePtr<T> fi = frame_at(frame_in.data(), 0, n_pad), fo = frame_at(frame_out.data(), 0, n_pad);
        for (int p = 0; p < cfg.n_particles; p++) p2g_body<T>(p, P, fi, fo, store_F, mat, grid_in.data());
        for (long long n = 0; n < n_nodes; n++) grid_fwd_body<T>(n, P, prims, s0, s1, grid_in.data(), grid_out.data(), false);
    }
};

template <class T>
int substep_fwd(const plb_config* c, const plb_primitive_desc* pd, double softness, const double* x, const double* v, const double* F,
                const double* C, const double* pose0, const double* pose1, double* xo, double* vo, double* Fo, double* Co,
                double* gin4, double* gout4) {
    Host<T> h(*c, pd, softness);
    h.pack(h.frame_in, x, v, F, C);
    Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    h.poses(pose0, pose1, s0, s1);
    h.forward_grid(s0, s1, true);
    FramePtr<T> fi = frame_at(h.frame_in.data(), 0, h.n_pad), fo = frame_at(h.frame_out.data(), 0, h.n_pad);
    for (int p = 0; p < c->n_particles; p++) g2p_body<T>(p, h.P, fi, fo, h.grid_out.data());
    h.unpack(h.frame_out, xo, vo, Fo, Co);
    for (long long n = 0; n < h.n_nodes; n++) {
        if (gin4) { gin4[n * 4] = h.grid_in[n].x; gin4[n * 4 + 1] = h.grid_in[n].y; gin4[n * 4 + 2] = h.grid_in[n].z; gin4[n * 4 + 3] = h.grid_in[n].w; }
        if (gout4) { gout4[n * 4] = h.grid_out[n].x; gout4[n * 4 + 1] = h.grid_out[n].y; gout4[n * 4 + 2] = h.grid_out[n].z; gout4[n * 4 + 3] = h.grid_out[n].w; }
    }
    return 0;
}

template <class T>
int substep_bwd(const plb_config* c, const plb_primitive_desc* pd, double softness, const double* x, const double* v, const double* F,
                const double* C, const double* pose0, const double* pose1, const double* gxn, const double* gvn, const double* gFn,
                const double* gCn, double* gx, double* gv, double* gF, double* gC, double* gpose0, double* gpose1) {
    Host<T> h(*c, pd, softness);
    h.pack(h.frame_in, x, v, F, C);
    h.pack(h.adj_next, gxn, gvn, gFn, gCn);
    Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    h.poses(pose0, pose1, s0, s1);
    h.forward_grid(s0, s1, false);
    FramePtr<T> fi = frame_at(h.frame_in.data(), 0, h.n_pad);
    FramePtr<T> an = frame_at(h.adj_next.data(), 0, h.n_pad), ac = frame_at(h.adj_cur.data(), 0, h.n_pad);
    for (int p = 0; p < c->n_particles; p++) g2p_bwd_body<T>(p, h.P, fi, an, ac, h.grid_out.data(), h.g_out.data());
    std::memset(gpose0, 0, sizeof(double) * 8 * c->n_primitives);
    std::memset(gpose1, 0, sizeof(double) * 8 * c->n_primitives);
    for (long long n = 0; n < h.n_nodes; n++) {
        PoseGrad<T> g0[PLB_MAX_PRIM], g1[PLB_MAX_PRIM];
        for (int k = 0; k < c->n_primitives; k++) { g0[k].clear(); g1[k].clear(); }
        unsigned touched = 0;
        grid_bwd_body<T>(n, h.P, h.prims, s0, s1, h.grid_in.data(), h.g_out.data(), h.g_in.data(), true, g0, g1, touched);
        for (int k = 0; k < c->n_primitives; k++) {
            if (!((touched >> k) & 1u)) continue;
            const PoseGrad<T>* gg[2] = {&g0[k], &g1[k]};
            double* dst[2] = {gpose0 + k * 8, gpose1 + k * 8};
            for (int w = 0; w < 2; w++) {
                dst[w][0] += gg[w]->pos.x; dst[w][1] += gg[w]->pos.y; dst[w][2] += gg[w]->pos.z;
                dst[w][3] += gg[w]->rot.w; dst[w][4] += gg[w]->rot.x; dst[w][5] += gg[w]->rot.y; dst[w][6] += gg[w]->rot.z;
                dst[w][7] += gg[w]->gap;
            }
        }
    }
    Material<T> mat{nullptr, nullptr, nullptr};
    for (int p = 0; p < c->n_particles; p++) p2g_bwd_body<T>(p, h.P, fi, an, ac, mat, h.g_in.data());
    h.unpack(h.adj_cur, gx, gv, gF, gC);
    return 0;
}

template <class T>
int loss_both(const plb_config* c, const plb_primitive_desc* pd, const double* x, const double* pose, const double* target,
              const double* target_sdf, double w_sdf, double w_density, double w_contact, int contact_all, double* out4,
              double* gx, double* gpose, int soft) {
    Host<T> h(*c, pd, 0.0);
    std::vector<double> zero3v(3 * c->n_particles, 0.0), zero9(9 * c->n_particles, 0.0);
    h.pack(h.frame_in, x, zero3v.data(), zero9.data(), zero9.data());
    FramePtr<T> fi = frame_at(h.frame_in.data(), 0, h.n_pad), ad = frame_at(h.adj_cur.data(), 0, h.n_pad);
    std::vector<T> gm(h.n_nodes, T(0)), tg(h.n_nodes), ts(h.n_nodes);
    for (long long n = 0; n < h.n_nodes; n++) { tg[n] = (T)target[n]; ts[n] = (T)target_sdf[n]; }
    for (int p = 0; p < c->n_particles; p++) loss_mass_body<T>(p, h.P, fi, gm.data());
    double density = 0, sdf = 0;
    for (long long n = 0; n < h.n_nodes; n++) { density += std::fabs((double)gm[n] - (double)tg[n]); sdf += (double)ts[n] * (double)gm[n]; }
    Pose<T> s0[PLB_MAX_PRIM];
    for (int k = 0; k < c->n_primitives; k++) s0[k] = load_pose<T>(pose + k * 8);
    double md[PLB_MAX_PRIM], nrm[PLB_MAX_PRIM], contact = 0;
    for (int k = 0; k < c->n_primitives; k++) {
        md[k] = soft ? 0.0 : 100000.0; nrm[k] = 0.0;
        if (!h.prims.s[k].movable) continue;
        for (int p = 0; p < c->n_particles; p++) {
            double d = (double)tmax(prim_sdf(h.prims.s[k], s0[k], load_x(fi, p)), T(0));
            if (soft) { double sw = 1.0 / (1.0 + d * d * 10000.0); nrm[k] += sw; md[k] += d * sw; }
            else if (d < md[k]) md[k] = d;
        }
        double v = soft ? md[k] / nrm[k] : md[k];
        contact += v * v;
    }
    out4[0] = contact * w_contact + density * w_density + sdf * w_sdf; out4[1] = contact; out4[2] = density; out4[3] = sdf;
    std::memset(gpose, 0, sizeof(double) * 8 * c->n_primitives);
    for (int p = 0; p < c->n_particles; p++) {
        PoseGrad<T> g[PLB_MAX_PRIM];
        for (int k = 0; k < c->n_primitives; k++) g[k].clear();
        unsigned touched = 0;
        loss_bwd_body<T>(p, h.P, fi, ad, gm.data(), tg.data(), ts.data(), (T)w_sdf, (T)w_density, (T)w_contact, h.prims, s0, md,
                         contact_all, g, touched, soft, nrm);
        for (int k = 0; k < c->n_primitives; k++) {
            if (!((touched >> k) & 1u)) continue;
            double* d = gpose + k * 8;
            d[0] += g[k].pos.x; d[1] += g[k].pos.y; d[2] += g[k].pos.z; d[3] += g[k].rot.w; d[4] += g[k].rot.x; d[5] += g[k].rot.y;
            d[6] += g[k].rot.z; d[7] += g[k].gap;
        }
    }
    std::vector<double> t3(3 * c->n_particles), t9(9 * c->n_particles), t9b(9 * c->n_particles);
    h.unpack(h.adj_cur, gx, t3.data(), t9.data(), t9b.data());
    return 0;
}

}  // namespace

extern "C" {

int emul_substep_fwd(int dtype, const plb_config* c, const plb_primitive_desc* pd, double softness, const double* x, const double* v,
                     const double* F, const double* C, const double* pose0, const double* pose1, double* xo, double* vo, double* Fo,
                     double* Co, double* gin4, double* gout4) {
    return dtype == PLB_F32 ? substep_fwd<float>(c, pd, softness, x, v, F, C, pose0, pose1, xo, vo, Fo, Co, gin4, gout4)
                            : substep_fwd<double>(c, pd, softness, x, v, F, C, pose0, pose1, xo, vo, Fo, Co, gin4, gout4);
}
int emul_substep_bwd(int dtype, const plb_config* c, const plb_primitive_desc* pd, double softness, const double* x, const double* v,
                     const double* F, const double* C, const double* pose0, const double* pose1, const double* gxn, const double* gvn,
                     const double* gFn, const double* gCn, double* gx, double* gv, double* gF, double* gC, double* gpose0, double* gpose1) {
    return dtype == PLB_F32
               ? substep_bwd<float>(c, pd, softness, x, v, F, C, pose0, pose1, gxn, gvn, gFn, gCn, gx, gv, gF, gC, gpose0, gpose1)
               : substep_bwd<double>(c, pd, softness, x, v, F, C, pose0, pose1, gxn, gvn, gFn, gCn, gx, gv, gF, gC, gpose0, gpose1);
}
int emul_loss(int dtype, const plb_config* c, const plb_primitive_desc* pd, const double* x, const double* pose, const double* target,
              const double* target_sdf, double w_sdf, double w_density, double w_contact, int contact_all, double* out4, double* gx,
              double* gpose, int soft) {
    return dtype == PLB_F32 ? loss_both<float>(c, pd, x, pose, target, target_sdf, w_sdf, w_density, w_contact, contact_all, out4, gx, gpose, soft)
                            : loss_both<double>(c, pd, x, pose, target, target_sdf, w_sdf, w_density, w_contact, contact_all, out4, gx, gpose, soft);
}
void emul_fk(const plb_primitive_desc* d, const double* st, const double* v, const double* w, double gv, double* out) {
    kin::fk_forward(make_kindesc(*d), st, v, w, gv, out);
}
void emul_fk_bwd(const plb_primitive_desc* d, const double* st, const double* v, const double* w, double gv, const double* gout,
                 double* gst, double* gvel, double* gw, double* ggv) {
    std::memset(gst, 0, 8 * sizeof(double)); std::memset(gvel, 0, 3 * sizeof(double)); std::memset(gw, 0, 3 * sizeof(double));
    *ggv = 0;
    kin::fk_backward(make_kindesc(*d), st, v, w, gv, gout, gst, gvel, gw, *ggv);
}
// kinematics reverse scan: whole episode at once (chunked = 0) or one env step at a time with the carried adjoint (chunked = 1).
// traj / vel / g: [n_steps * S + 1][PLB_MAX_PRIM][8]; inject (optional): [n_steps][n_prim][8] added to the carry after each
// step's scan (policy observation adjoint).  out: [n_steps][action_total].  Returns action_total, or -1 on an order error.
int emul_action_grad(const plb_primitive_desc* pd, int n_prim, const double* traj, const double* vel, const double* g_dev, int n_steps,
                     int S, int chunked, const double* inject, double* out) {
    std::vector<kin::Desc> kd;
    std::vector<int> off;
    int total = 0;
    for (int k = 0; k < n_prim; k++) { kd.push_back(make_kindesc(pd[k])); off.push_back(total); total += pd[k].action_dim; }
    const size_t row = (size_t)PLB_MAX_PRIM * 8;
    const int nf = n_steps * S;
    std::memset(out, 0, sizeof(double) * (size_t)n_steps * total);
    if (!chunked) {
        std::vector<double> g(g_dev, g_dev + (size_t)(nf + 1) * row);
        if (inject)      // the observation adjoint of step t lands on the pose of frame t*S
            for (int t = 0; t < n_steps; t++)
                for (int k = 0; k < n_prim; k++)
                    for (int i = 0; i < 8; i++) g[(size_t)t * S * row + (size_t)k * 8 + i] += inject[((size_t)t * n_prim + k) * 8 + i];
        kin::action_grad_scan(kd.data(), n_prim, PLB_MAX_PRIM, traj, vel, g.data(), 0, 0, nf, S, off.data(), total, out, 0);
        return total;
    }
    kin::ScanCarry carry;
    std::vector<double> scratch((size_t)(S + 1) * row);
    for (int t = n_steps - 1; t >= 0; t--) {
        if (!kin::action_grad_step(kd.data(), n_prim, traj, vel, g_dev + (size_t)t * S * row, t, S, off.data(), total, carry,
                                   out + (size_t)t * total, scratch.data()))
            return -1;
        if (inject)
            for (int k = 0; k < n_prim; k++)
                for (int i = 0; i < 8; i++) carry.v[(size_t)k * 8 + i] += inject[((size_t)t * n_prim + k) * 8 + i];
    }
    return total;
}
void emul_svd(int dtype, const double* F9, double* U9, double* s3, double* V9) {
    if (dtype == PLB_F32) {
        M3<float> F, U, V; V3<float> s;
        for (int i = 0; i < 9; i++) F.m[i / 3][i % 3] = (float)F9[i];
        svd3(F, U, s, V);
        for (int i = 0; i < 9; i++) { U9[i] = U.m[i / 3][i % 3]; V9[i] = V.m[i / 3][i % 3]; }
        s3[0] = s.x; s3[1] = s.y; s3[2] = s.z;
    } else {
        M3<double> F, U, V; V3<double> s;
        for (int i = 0; i < 9; i++) F.m[i / 3][i % 3] = F9[i];
        svd3(F, U, s, V);
        for (int i = 0; i < 9; i++) { U9[i] = U.m[i / 3][i % 3]; V9[i] = V.m[i / 3][i % 3]; }
        s3[0] = s.x; s3[1] = s.y; s3[2] = s.z;
    }
}
// same with a starting V (the decomposition of a nearby matrix); returns nothing else
void emul_svd_warm(int dtype, const double* F9, const double* W9, double* U9, double* s3, double* V9) {
    if (dtype == PLB_F32) {
        M3<float> F, U, V, W; V3<float> s;
        for (int i = 0; i < 9; i++) { F.m[i / 3][i % 3] = (float)F9[i]; W.m[i / 3][i % 3] = (float)W9[i]; }
        svd3(F, U, s, V, &W);
        for (int i = 0; i < 9; i++) { U9[i] = U.m[i / 3][i % 3]; V9[i] = V.m[i / 3][i % 3]; }
        s3[0] = s.x; s3[1] = s.y; s3[2] = s.z;
    } else {
        M3<double> F, U, V, W; V3<double> s;
        for (int i = 0; i < 9; i++) { F.m[i / 3][i % 3] = F9[i]; W.m[i / 3][i % 3] = W9[i]; }
        svd3(F, U, s, V, &W);
        for (int i = 0; i < 9; i++) { U9[i] = U.m[i / 3][i % 3]; V9[i] = V.m[i / 3][i % 3]; }
        s3[0] = s.x; s3[1] = s.y; s3[2] = s.z;
    }
}

}  // extern "C"
