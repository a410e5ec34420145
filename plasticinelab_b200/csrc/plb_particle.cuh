// Per-particle math of the MLS-MPM substep and its hand-derived adjoint.
//
// Forward semantics follow plb/engine/mpm_simulator.py:
//   compute_F_tmp :82-85, svd :87-90, compute_von_mises :124-141, norm :153-155, p2g :157-184, g2p :223-242.
// The adjoint follows the order of substep_grad (:260-278): g2p.grad, (grid), p2g.grad, svd_grad, compute_F_tmp.grad,
// with the reference's max/min gradient routing (Taichi autodiff) and its SVD adjoint.
#pragma once
#include "plb_svd.cuh"

namespace plb {

// Constants of one simulator instance, in the kernel's scalar type.
template <class T> struct SimConst {
    T dx, inv_dx, dt, p_vol, p_mass;
    T stress_scale;       // -dt * p_vol * 4 * inv_dx^2
    T x_hi;               // 1 - 3 dx
    T grav_dv[3];         // dt * gravity * 30
    T ground_friction;
    int n_grid;
    int n_particles;
    int n_prim;
    T mu, lam, yield_stress;   // uniform material (used when no per-particle arrays are given)
};

// Quadratic B-spline stencil of one particle: base node and weights w[a][d] (a = offset 0..2, d = axis).
template <class T> struct Stencil {
    int b[3];
    V3<T> fx;
    T w[3][3];
};

template <class T> PLB_HD Stencil<T> make_stencil(V3<T> x, T inv_dx) {
    Stencil<T> s;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        T xs = x[d] * inv_dx;
        int b = (int)(xs - T(0.5));          // C truncation, like ti .cast(int)
        T f = xs - (T)b;
        s.b[d] = b;
        s.fx[d] = f;
        s.w[0][d] = T(0.5) * (T(1.5) - f) * (T(1.5) - f);
        s.w[1][d] = T(0.75) - (f - T(1)) * (f - T(1));
        s.w[2][d] = T(0.5) * (f - T(0.5)) * (f - T(0.5));
    }
    return s;
}

// d w[a][d] / d fx[d]
template <class T> PLB_HD T dweight(int a, T f) {
    return a == 0 ? -(T(1.5) - f) : (a == 1 ? T(-2) * (f - T(1)) : (f - T(0.5)));
}

// Chain gw[a][d] (adjoint of the 1-D weights) and gfx_direct (adjoint reaching fx through dpos) to x.
template <class T> PLB_HD V3<T> stencil_backward(const Stencil<T>& s, const T gw[3][3], V3<T> gfx, T inv_dx) {
    V3<T> gx;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        T g = gfx[d];
#pragma unroll
        for (int a = 0; a < 3; a++) g += gw[a][d] * dweight<T>(a, s.fx[d]);
        gx[d] = g * inv_dx;
    }
    return gx;
}

// ------------------------------------------------------------------------------------------------
// P2G, particle part
// ------------------------------------------------------------------------------------------------
template <class T> struct P2GState {      // everything the adjoint needs from the forward
    M3<T> F_tmp, U, V, new_F, M;          // M = new_F - U V^T
    V3<T> sig, eps_hat, e;                // sig raw; e = exp(eps') (yield branch)
    T eps_norm, J, c;                     // c = yield / (2 mu)
    bool yield;
};

// SVD of F_tmp of one particle and substep.  The forward pass can keep it per frame (21 scalars per particle: 5 Vec4 planes +
// 1 scalar plane, `SvdPtr`) so that the backward pass loads it instead of re-running the Jacobi iteration (~500 of the fused
// backward kernel's ~4400 instructions per warp); HBM has the headroom (the particle kernels use 2-19 % of it).
template <class T> struct SvdRec { M3<T> U, V; V3<T> sig; };
template <class T> struct SvdPtr { Vec4<T>* q[5]; T* s; };
constexpr int kSvdScalars = 21;
template <class T> PLB_HD SvdPtr<T> svd_at(T* base, long long slot, long long n_pad) {
    T* b = base + slot * kSvdScalars * n_pad;
    SvdPtr<T> r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.q[i] = reinterpret_cast<Vec4<T>*>(b + 4 * i * n_pad);
    r.s = b + 20 * n_pad;
    return r;
}
template <class T> PLB_HD void store_svd(const SvdPtr<T>& f, int p, const SvdRec<T>& r) {
    f.q[0][p] = mk4<T>(r.U.m[0][0], r.U.m[0][1], r.U.m[0][2], r.U.m[1][0]);
    f.q[1][p] = mk4<T>(r.U.m[1][1], r.U.m[1][2], r.U.m[2][0], r.U.m[2][1]);
    f.q[2][p] = mk4<T>(r.V.m[0][0], r.V.m[0][1], r.V.m[0][2], r.V.m[1][0]);
    f.q[3][p] = mk4<T>(r.V.m[1][1], r.V.m[1][2], r.V.m[2][0], r.V.m[2][1]);
    f.q[4][p] = mk4<T>(r.U.m[2][2], r.V.m[2][2], r.sig.x, r.sig.y);
    f.s[p] = r.sig.z;
}
// V alone (3 of the 6 planes): warm start of the next substep's iteration
template <class T> PLB_HD M3<T> load_svd_V(const SvdPtr<T>& f, int p) {
    Vec4<T> c = f.q[2][p], d = f.q[3][p], e = f.q[4][p];
    M3<T> V;
    V.m[0][0] = c.x; V.m[0][1] = c.y; V.m[0][2] = c.z; V.m[1][0] = c.w;
    V.m[1][1] = d.x; V.m[1][2] = d.y; V.m[2][0] = d.z; V.m[2][1] = d.w;
    V.m[2][2] = e.y;
    return V;
}
template <class T> PLB_HD SvdRec<T> load_svd(const SvdPtr<T>& f, int p) {
    Vec4<T> a = f.q[0][p], b = f.q[1][p], c = f.q[2][p], d = f.q[3][p], e = f.q[4][p];
    SvdRec<T> r;
    r.U.m[0][0] = a.x; r.U.m[0][1] = a.y; r.U.m[0][2] = a.z; r.U.m[1][0] = a.w;
    r.U.m[1][1] = b.x; r.U.m[1][2] = b.y; r.U.m[2][0] = b.z; r.U.m[2][1] = b.w;
    r.V.m[0][0] = c.x; r.V.m[0][1] = c.y; r.V.m[0][2] = c.z; r.V.m[1][0] = c.w;
    r.V.m[1][1] = d.x; r.V.m[1][2] = d.y; r.V.m[2][0] = d.z; r.V.m[2][1] = d.w;
    r.U.m[2][2] = e.x; r.V.m[2][2] = e.y;
    r.sig = mk3<T>(e.z, e.w, f.s[p]);
    return r;
}

// Forward: returns new_F (= F[f+1]) and the APIC affine matrix (stress + p_mass C).
// kSvdGiven: `svd` holds the decomposition of F_tmp (loaded from the store); otherwise it is computed and, if svd != nullptr,
// returned through it.  warmV (optional, with !kSvdGiven): V of this particle's previous substep, starting point of the iteration.
template <class T, bool kSvdGiven = false>
PLB_HD void p2g_particle(const SimConst<T>& P, const M3<T>& C, const M3<T>& F, T mu, T lam, T ys,
                         M3<T>& new_F, M3<T>& affine, P2GState<T>* keep = nullptr, SvdRec<T>* svd = nullptr, const M3<T>* warmV = nullptr) {
    M3<T> A = identM<T>() + P.dt * C;
    M3<T> F_tmp = mm(A, F);
    M3<T> U, V;
    V3<T> sig;
    if (kSvdGiven) {
        U = svd->U; V = svd->V; sig = svd->sig;
    } else {
        svd3(F_tmp, U, sig, V, warmV);
        if (svd) { svd->U = U; svd->V = V; svd->sig = sig; }
    }
    // compute_von_mises
    V3<T> sc = mk3<T>(tmax(sig.x, T(0.05)), tmax(sig.y, T(0.05)), tmax(sig.z, T(0.05)));
    V3<T> eps = mk3<T>(plb_log(sc.x), plb_log(sc.y), plb_log(sc.z));
    T mean = (eps.x + eps.y + eps.z) * T(1.0 / 3.0);
    V3<T> eh = mk3<T>(eps.x - mean, eps.y - mean, eps.z - mean);
    T n = plb_sqrt(dot(eh, eh) + T(1e-8));
    T c = ys * plb_rcp_nr(T(2) * mu);
    T dgamma = n - c;
    bool yield = dgamma > T(0);
    V3<T> e = zero3<T>();
    if (yield) {
        T k = dgamma / n;
        e = mk3<T>(plb_exp(eps.x - k * eh.x), plb_exp(eps.y - k * eh.y), plb_exp(eps.z - k * eh.z));
        M3<T> UE;
#pragma unroll
        for (int i = 0; i < 3; i++) { UE.m[i][0] = U.m[i][0] * e.x; UE.m[i][1] = U.m[i][1] * e.y; UE.m[i][2] = U.m[i][2] * e.z; }
        new_F = mmT(UE, V);
    } else {
        new_F = F_tmp;
    }
    T J = det(new_F);
    M3<T> r = mmT(U, V);
    M3<T> M = new_F - r;
    M3<T> stress = (T(2) * mu) * mmT(M, new_F);
    T lj = lam * J * (J - T(1));
    stress.m[0][0] += lj; stress.m[1][1] += lj; stress.m[2][2] += lj;
    affine = P.stress_scale * stress + P.p_mass * C;
    if (keep) {
        keep->F_tmp = F_tmp; keep->U = U; keep->V = V; keep->new_F = new_F; keep->M = M;
        keep->sig = sig; keep->eps_hat = eh; keep->e = e; keep->eps_norm = n; keep->J = J; keep->c = c;
        keep->yield = yield;
    }
}

// Adjoint of the particle part.  Inputs: g_affine (adjoint of `affine`), gF_next (adjoint of F[f+1]).
// Outputs: gC, gF (adjoints of C[f], F[f]); the caller adds the stencil terms to gx and gv itself.
template <class T>
PLB_HD void p2g_particle_backward(const SimConst<T>& P, const M3<T>& C, const M3<T>& F, T mu, T lam,
                                  const P2GState<T>& k, const M3<T>& g_affine, const M3<T>& gF_next,
                                  M3<T>& gC, M3<T>& gF) {
    gC = P.p_mass * g_affine;
    M3<T> S = P.stress_scale * g_affine;               // adjoint of `stress`
    // stress = 2 mu M new_F^T + I lam J (J - 1)
    M3<T> gM = (T(2) * mu) * mm(S, k.new_F);
    M3<T> gNF = gM + (T(2) * mu) * mTm(S, k.M) + gF_next;
    T gJ = lam * (T(2) * k.J - T(1)) * (S.m[0][0] + S.m[1][1] + S.m[2][2]);
    gNF += gJ * cofactor(k.new_F);
    // r = U V^T,  M = new_F - r  ->  g(r) = -gM
    M3<T> gU = (T(-1)) * mm(gM, k.V);
    M3<T> gV = (T(-1)) * mTm(gM, k.U);
    V3<T> gsig = zero3<T>();
    M3<T> gFtmp;
    if (k.yield) {
        // new_F = U E V^T
        M3<T> VE, UE;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            VE.m[i][0] = k.V.m[i][0] * k.e.x; VE.m[i][1] = k.V.m[i][1] * k.e.y; VE.m[i][2] = k.V.m[i][2] * k.e.z;
            UE.m[i][0] = k.U.m[i][0] * k.e.x; UE.m[i][1] = k.U.m[i][1] * k.e.y; UE.m[i][2] = k.U.m[i][2] * k.e.z;
        }
        gU += mm(gNF, VE);
        gV += mTm(gNF, UE);
        M3<T> UtGV = mm(mTm(k.U, gNF), k.V);
        V3<T> geps1 = mk3<T>(UtGV.m[0][0] * k.e.x, UtGV.m[1][1] * k.e.y, UtGV.m[2][2] * k.e.z);   // adjoint of eps'
        // eps' = eps - kk * eps_hat, kk = dgamma / n = 1 - c / n
        T n = k.eps_norm;
        T kk = T(1) - k.c / n;
        V3<T> geps = geps1;
        V3<T> geh = (-kk) * geps1;
        T gk = -dot(geps1, k.eps_hat);
        T gn = gk * k.c / (n * n);
        geh += (gn / n) * k.eps_hat;
        T gmean = (geh.x + geh.y + geh.z) / T(3);
        geps += mk3<T>(geh.x - gmean, geh.y - gmean, geh.z - gmean);
        // eps = log(max(sig, 0.05))
#pragma unroll
        for (int i = 0; i < 3; i++) {
            T s = k.sig[i];
            gsig[i] = (T(0.05) < s) ? geps[i] / s : T(0);
        }
        gFtmp = zeroM<T>();
    } else {
        gFtmp = gNF;
    }
    gFtmp += svd3_backward(gU, gsig, gV, k.U, k.sig, k.V);
    // F_tmp = (I + dt C) F
    gC += P.dt * mmT(gFtmp, F);
    M3<T> A = identM<T>() + P.dt * C;
    gF = mTm(A, gFtmp);
}

// ------------------------------------------------------------------------------------------------
// G2P, particle part (after the 27-node gather produced new_v and new_C)
// ------------------------------------------------------------------------------------------------
template <class T> PLB_HD V3<T> advect(const SimConst<T>& P, V3<T> x, V3<T> new_v) {
    V3<T> r;
#pragma unroll
    for (int d = 0; d < 3; d++) r[d] = tmax(tmin(x[d] + P.dt * new_v[d], P.x_hi), T(0));
    return r;
}

// adjoint of advect: returns gy (the part of gx_next that passes the two clamps)
template <class T> PLB_HD V3<T> advect_backward(const SimConst<T>& P, V3<T> x, V3<T> new_v, V3<T> gx_next) {
    V3<T> r;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        T y = x[d] + P.dt * new_v[d];
        bool pass_min = y < P.x_hi;                      // tmin(y, hi): to y iff y < hi
        T z = pass_min ? y : P.x_hi;
        bool pass_max = T(0) < z;                        // tmax(z, 0): to z iff 0 < z
        r[d] = (pass_min && pass_max) ? gx_next[d] : T(0);
    }
    return r;
}

// same, from the position G2P stored: x' = max(min(y, hi), 0) lies strictly inside (0, hi) iff both clamps passed y through
template <class T> PLB_HD V3<T> advect_backward_stored(const SimConst<T>& P, V3<T> x_next, V3<T> gx_next) {
    V3<T> r;
#pragma unroll
    for (int d = 0; d < 3; d++) r[d] = (T(0) < x_next[d] && x_next[d] < P.x_hi) ? gx_next[d] : T(0);
    return r;
}

}  // namespace plb
