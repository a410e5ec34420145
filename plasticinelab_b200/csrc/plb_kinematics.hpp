// Host-side (float64) primitive kinematics and their adjoint.
//
// The primitive trajectory depends on the actions only, never on the particles, so it is integrated on the host
// (a few hundred flops per substep) and uploaded once per env step; the device kernels read poses f and f+1.
// Reference: forward_kinematics  plb/engine/primitive/primive_base.py:117-121   (base: q' = w2quat(w) * q)
//                                plb/engine/primitive/primitives.py:66-80        (RollingPin)
//                                plb/engine/primitive/primitives.py:94-98        (Chopsticks: q' = q * w2quat(w), gap)
//            set_velocity        primive_base.py:184-192, primitives.py:101-109
//            qmul / w2quat       plb/engine/primitive/utils.py:19-41
// The adjoints mirror Taichi's reverse mode of those kernels (max/min route by strict comparison).
#pragma once
#include <cmath>
#include <cstring>
#include "plb_primitives.cuh"

namespace plb {
namespace kin {

struct Desc {
    int type;
    double lower[3], upper[3];
    int action_dim;
    double action_scale[8];
    double minimal_gap;
};

inline void qmul(const double q[4], const double r[4], double out[4], double* raw = nullptr) {
    double n[4];
    n[0] = r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3];
    n[1] = r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2];
    n[2] = r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1];
    n[3] = r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0];
    double l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3]);
    for (int i = 0; i < 4; i++) out[i] = n[i] / l;
    if (raw) std::memcpy(raw, n, sizeof(n));
}

inline void qmul_bwd(const double q[4], const double r[4], const double gout[4], double gq[4], double gr[4]) {
    double n[4], o[4];
    qmul(q, r, o, n);
    double l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3]);
    double d = o[0] * gout[0] + o[1] * gout[1] + o[2] * gout[2] + o[3] * gout[3];
    double g[4];
    for (int i = 0; i < 4; i++) g[i] = (gout[i] - o[i] * d) / l;
    // n0 = r0q0 - r1q1 - r2q2 - r3q3
    gq[0] += g[0] * r[0]; gq[1] -= g[0] * r[1]; gq[2] -= g[0] * r[2]; gq[3] -= g[0] * r[3];
    gr[0] += g[0] * q[0]; gr[1] -= g[0] * q[1]; gr[2] -= g[0] * q[2]; gr[3] -= g[0] * q[3];
    // n1 = r0q1 + r1q0 - r2q3 + r3q2
    gq[1] += g[1] * r[0]; gq[0] += g[1] * r[1]; gq[3] -= g[1] * r[2]; gq[2] += g[1] * r[3];
    gr[0] += g[1] * q[1]; gr[1] += g[1] * q[0]; gr[2] -= g[1] * q[3]; gr[3] += g[1] * q[2];
    // n2 = r0q2 + r1q3 + r2q0 - r3q1
    gq[2] += g[2] * r[0]; gq[3] += g[2] * r[1]; gq[0] += g[2] * r[2]; gq[1] -= g[2] * r[3];
    gr[0] += g[2] * q[2]; gr[1] += g[2] * q[3]; gr[2] += g[2] * q[0]; gr[3] -= g[2] * q[1];
    // n3 = r0q3 - r1q2 + r2q1 + r3q0
    gq[3] += g[3] * r[0]; gq[2] -= g[3] * r[1]; gq[1] += g[3] * r[2]; gq[0] += g[3] * r[3];
    gr[0] += g[3] * q[3]; gr[1] -= g[3] * q[2]; gr[2] += g[3] * q[1]; gr[3] += g[3] * q[0];
}

inline void w2quat(const double w[3], double out[4]) {
    double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    out[0] = 1.0; out[1] = out[2] = out[3] = 0.0;
    if (n > 1e-9) {
        double s = std::sin(n / 2);
        out[0] = std::cos(n / 2);
        for (int i = 0; i < 3; i++) out[1 + i] = w[i] / n * s;
    }
}

// For |w| <= 1e-9 no gradient reaches w (Taichi would produce 0 * inf there; defined as 0).
inline void w2quat_bwd(const double w[3], const double g[4], double gw[3]) {
    double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (!(n > 1e-9)) return;
    double s = std::sin(n / 2), c = std::cos(n / 2);
    double gn = -0.5 * s * g[0];
    double gs = 0, gu[3];
    for (int i = 0; i < 3; i++) { gu[i] = g[1 + i] * s; gs += g[1 + i] * w[i] / n; }
    gn += 0.5 * c * gs;
    double dotuw = 0;
    for (int i = 0; i < 3; i++) dotuw += gu[i] * w[i];
    gn -= dotuw / (n * n);
    for (int i = 0; i < 3; i++) gw[i] += gu[i] / n + gn * w[i] / n;
}

inline double clampf(double x, double lo, double hi, double& dpass) {
    // max(min(x, hi), lo) with Taichi gradient routing
    bool pmin = x < hi;
    double z = pmin ? x : hi;
    bool pmax = lo < z;
    dpass = (pmin && pmax) ? 1.0 : 0.0;
    return pmax ? z : lo;
}

// state: pos(3) rot(4) gap ; v(3) w(3) gv : per-substep velocities
inline void fk_forward(const Desc& d, const double st[8], const double v[3], const double w[3], double gv, double out[8]) {
    double dp;
    if (d.type == PRIM_ROLLINGPIN) {
        double dw = v[0], dth = v[1], dy = v[2];
        Q4<double> rot{st[3], st[4], st[5], st[6]};
        V3<double> ydir = qrot(rot, mk3<double>(0.0, -1.0, 0.0));
        V3<double> xdir = cross(mk3<double>(0.0, 1.0, 0.0), ydir) * (dw * 0.03);
        xdir.y = dy;
        double a[3] = {0.0, -dth, 0.0}, b[3] = {0.0, dw, 0.0}, qa[4], qb[4], tmp[4];
        w2quat(a, qa); w2quat(b, qb);
        qmul(st + 3, qb, tmp);            // qmul(rotation, w2quat(0,dw,0))
        qmul(qa, tmp, out + 3);           // qmul(w2quat(0,-dth,0), .)
        for (int i = 0; i < 3; i++) out[i] = clampf(st[i] + xdir[i], d.lower[i], d.upper[i], dp);
        out[7] = st[7];
        return;
    }
    for (int i = 0; i < 3; i++) out[i] = clampf(st[i] + v[i], d.lower[i], d.upper[i], dp);
    double qw[4];
    w2quat(w, qw);
    if (d.type == PRIM_CHOPSTICKS) {
        qmul(st + 3, qw, out + 3);        // right multiply
        double z = st[7] - gv;
        out[7] = (d.minimal_gap < z) ? z : d.minimal_gap;     // tmax(z, minimal_gap)
    } else {
        qmul(qw, st + 3, out + 3);        // left multiply
        out[7] = st[7];
    }
}

// adjoint: gout(8) -> gst(8) +=, gvel(3) +=, gw(3) +=, ggv +=
inline void fk_backward(const Desc& d, const double st[8], const double v[3], const double w[3], double gv,
                        const double gout[8], double gst[8], double gvel[3], double gw[3], double& ggv) {
    double dp;
    if (d.type == PRIM_ROLLINGPIN) {
        double dw = v[0], dth = v[1];
        Q4<double> rot{st[3], st[4], st[5], st[6]};
        V3<double> e1 = mk3<double>(0.0, 1.0, 0.0), my = mk3<double>(0.0, -1.0, 0.0);
        V3<double> ydir = qrot(rot, my);
        V3<double> cr = cross(e1, ydir);
        V3<double> xdir = cr * (dw * 0.03);
        xdir.y = v[2];
        V3<double> gx;
        for (int i = 0; i < 3; i++) { clampf(st[i] + xdir[i], d.lower[i], d.upper[i], dp); gx[i] = gout[i] * dp; gst[i] += gx[i]; }
        // xdir = cross(e1, ydir) * dw * 0.03 with [1] overwritten by dy
        gvel[2] += gx.y;
        V3<double> gxd = mk3<double>(gx.x, 0.0, gx.z);
        gvel[0] += 0.03 * dot(gxd, cr);
        V3<double> gcr = (dw * 0.03) * gxd;
        V3<double> gyd = cross(gcr, e1);            // d/d ydir of gcr . (e1 x ydir) = gcr x e1
        Q4<double> grot{0, 0, 0, 0};
        qrot_bwd_q(rot, my, gyd, grot);
        // rotation chain
        double a[3] = {0.0, -dth, 0.0}, b[3] = {0.0, dw, 0.0}, qa[4], qb[4], tmp[4];
        w2quat(a, qa); w2quat(b, qb);
        qmul(st + 3, qb, tmp);
        double gqa[4] = {0, 0, 0, 0}, gtmp[4] = {0, 0, 0, 0}, gq[4] = {0, 0, 0, 0}, gqb[4] = {0, 0, 0, 0};
        qmul_bwd(qa, tmp, gout + 3, gqa, gtmp);
        qmul_bwd(st + 3, qb, gtmp, gq, gqb);
        double ga[3] = {0, 0, 0}, gb[3] = {0, 0, 0};
        w2quat_bwd(a, gqa, ga);
        w2quat_bwd(b, gqb, gb);
        gvel[1] -= ga[1];
        gvel[0] += gb[1];
        gst[3] += gq[0] + grot.w; gst[4] += gq[1] + grot.x; gst[5] += gq[2] + grot.y; gst[6] += gq[3] + grot.z;
        gst[7] += gout[7];
        return;
    }
    for (int i = 0; i < 3; i++) { clampf(st[i] + v[i], d.lower[i], d.upper[i], dp); gst[i] += gout[i] * dp; gvel[i] += gout[i] * dp; }
    double qw[4], gqw[4] = {0, 0, 0, 0}, gq[4] = {0, 0, 0, 0};
    w2quat(w, qw);
    if (d.type == PRIM_CHOPSTICKS) {
        qmul_bwd(st + 3, qw, gout + 3, gq, gqw);
        double z = st[7] - gv;
        if (d.minimal_gap < z) { gst[7] += gout[7]; ggv -= gout[7]; }
    } else {
        qmul_bwd(qw, st + 3, gout + 3, gqw, gq);
        gst[7] += gout[7];
    }
    w2quat_bwd(w, gqw, gw);
    for (int i = 0; i < 4; i++) gst[3 + i] += gq[i];
}

// Reverse scan of the pose adjoints through forward_kinematics.grad and set_velocity.grad for the frames f_hi-1 .. f_lo
// (Primitive.forward_kinematics.grad + set_velocity.grad replayed by the tape, primive_base.py:117-121,184-192).
//   g      pose adjoints [frame - g_base][max_prim][8]; g[f+1] is consumed, g[f] accumulates (in place)
//   traj   poses [frame][max_prim][8], vel per-substep velocities [frame][max_prim][8] = v(3) w(3) gap-velocity pad
//   out    action gradients [step - out_base][action_total], step = f / S; ACCUMULATED into
// Used for the whole episode at once (plb_get_action_grad) and one env step at a time (plb_action_grad_step, policy path).
inline void action_grad_scan(const Desc* descs, int n_prim, int max_prim, const double* traj, const double* vel, double* g, int g_base,
                             int f_lo, int f_hi, int S, const int* action_off, int action_total, double* out, int out_base) {
    for (int f = f_hi - 1; f >= f_lo; f--) {
        const int step = f / S;
        for (int k = n_prim - 1; k >= 0; k--) {
            const Desc& d = descs[k];
            const double* vv = vel + ((size_t)f * max_prim + k) * 8;
            const double* gnext = g + ((size_t)(f + 1 - g_base) * max_prim + k) * 8;
            double* gcur = g + ((size_t)(f - g_base) * max_prim + k) * 8;
            double gvel[3] = {0, 0, 0}, gw[3] = {0, 0, 0}, ggv = 0;
            fk_backward(d, traj + ((size_t)f * max_prim + k) * 8, vv, vv + 3, vv[6], gnext, gcur, gvel, gw, ggv);
            if (d.action_dim == 0) continue;
            double* o = out + (size_t)(step - out_base) * action_total + action_off[k];
            for (int i = 0; i < 3; i++) o[i] += gvel[i] * d.action_scale[i] / S;
            if (d.action_dim > 3) for (int i = 0; i < 3; i++) o[3 + i] += gw[i] * d.action_scale[3 + i] / S;
            if (d.type == PRIM_CHOPSTICKS) o[6] += ggv * d.action_scale[6] / S;
        }
    }
}

// One env step of that scan for state-feedback policies (the action of step t depends on the state at frame t*S, so its
// gradient is needed while the backward sweep stands there).  `carry` holds the adjoint that flows into the pose of frame
// (step+1)*S from everything after it; `dev` = the accumulated pose adjoints of the frames step*S .. (step+1)*S as the device
// holds them now ([S+1][PLB_MAX_PRIM][8]; the row of frame step*S is still incomplete, which is why only the part the scan
// ADDS to it is carried on).  Steps must come in descending order.  out[action_total] is overwritten.
struct ScanCarry {
    double v[PLB_MAX_PRIM * 8];
    int frame;                       // frame the carry belongs to; -1: nothing carried yet
    ScanCarry() { reset(); }
    void reset() { std::memset(v, 0, sizeof(v)); frame = -1; }
};
inline bool action_grad_step(const Desc* descs, int n_prim, const double* traj, const double* vel, const double* dev, int step, int S,
                             const int* action_off, int action_total, ScanCarry& carry, double* out, double* scratch /*[(S+1)*PLB_MAX_PRIM*8]*/) {
    const int f_lo = step * S, f_hi = (step + 1) * S;
    if (carry.frame >= 0 && carry.frame != f_hi) return false;
    const size_t row = (size_t)PLB_MAX_PRIM * 8;
    std::memcpy(scratch, dev, (size_t)(S + 1) * row * sizeof(double));
    if (carry.frame == f_hi) for (size_t i = 0; i < row; i++) scratch[(size_t)S * row + i] += carry.v[i];
    for (int i = 0; i < action_total; i++) out[i] = 0.0;
    action_grad_scan(descs, n_prim, PLB_MAX_PRIM, traj, vel, scratch, f_lo, f_lo, f_hi, S, action_off, action_total, out, step);
    for (size_t i = 0; i < row; i++) carry.v[i] = scratch[i] - dev[i];
    carry.frame = f_lo;
    return true;
}

}  // namespace kin
}  // namespace plb
