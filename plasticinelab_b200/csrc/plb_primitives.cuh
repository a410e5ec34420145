// Rigid analytic-SDF manipulators: signed distance, normal, contact response and their adjoints.
//
// Restated from
//   plb/engine/primitive/primive_base.py:57-115 (sdf / normal / collider_v / collide)
//   plb/engine/primitive/primitives.py:8-257    (shape SDFs; `length` with eps 1e-14)
//   plb/engine/primitive/utils.py:3-47          (`length` with eps 1e-8, qrot, inv_trans)
// The adjoints are hand-derived; max/min route gradients by strict comparison as Taichi's autodiff does.
#pragma once
#include "plb_types.cuh"

namespace plb {

enum PrimType : int { PRIM_SPHERE = 0, PRIM_CAPSULE = 1, PRIM_ROLLINGPIN = 2, PRIM_CHOPSTICKS = 3,
                      PRIM_CYLINDER = 4, PRIM_TORUS = 5, PRIM_BOX = 6 };

constexpr int PLB_MAX_PRIM = 8;
constexpr int PLB_POSE_DIM = 8;          // position(3), rotation(4), gap(1)

template <class T> struct PrimStatic {
    int type;
    int movable;          // action_dim > 0 (takes part in the contact loss)
    T p[4];               // sphere: radius | capsule family: h, r | cylinder: h, r | torus: tx, ty | box: size xyz
    T friction;
    T softness;
};

template <class T> struct Pose { V3<T> pos; Q4<T> rot; T gap; };

template <class T> struct PoseGrad {
    V3<T> pos; Q4<T> rot; T gap;
    PLB_HD void clear() { pos = zero3<T>(); rot.w = rot.x = rot.y = rot.z = T(0); gap = T(0); }
};

template <class T> PLB_HD Pose<T> load_pose(const double* s) {
    Pose<T> p;
    p.pos = mk3<T>((T)s[0], (T)s[1], (T)s[2]);
    p.rot.w = (T)s[3]; p.rot.x = (T)s[4]; p.rot.y = (T)s[5]; p.rot.z = (T)s[6];
    p.gap = (T)s[7];
    return p;
}

// ---------------------------------------------------------------- quaternion rotation and its adjoints
template <class T> PLB_HD V3<T> qrot(Q4<T> q, V3<T> v) {
    V3<T> u = mk3<T>(q.x, q.y, q.z);
    V3<T> uv = cross(u, v);
    V3<T> uuv = cross(u, uv);
    return v + T(2) * (q.w * uv + uuv);
}
// adjoint wrt v
template <class T> PLB_HD V3<T> qrot_bwd_v(Q4<T> q, V3<T> g) {
    V3<T> u = mk3<T>(q.x, q.y, q.z);
    V3<T> ug = cross(u, g);
    return g + T(2) * (cross(u, ug) - q.w * ug);
}
// adjoint wrt q (accumulates)
template <class T> PLB_HD void qrot_bwd_q(Q4<T> q, V3<T> v, V3<T> g, Q4<T>& gq) {
    V3<T> u = mk3<T>(q.x, q.y, q.z);
    gq.w += T(2) * dot(g, cross(u, v));
    V3<T> gu = T(2) * (q.w * cross(v, g) + dot(u, v) * g + dot(g, u) * v - (T(2) * dot(g, v)) * u);
    gq.x += gu.x; gq.y += gu.y; gq.z += gu.z;
}

// inv_trans(p, position, rotation) = qrot(normalize(conj(rotation)), p - position)
template <class T> PLB_HD V3<T> inv_trans(V3<T> p, const Pose<T>& s) {
    T n = plb_sqrt(s.rot.w * s.rot.w + s.rot.x * s.rot.x + s.rot.y * s.rot.y + s.rot.z * s.rot.z);
    Q4<T> qi; qi.w = s.rot.w / n; qi.x = -s.rot.x / n; qi.y = -s.rot.y / n; qi.z = -s.rot.z / n;
    return qrot(qi, p - s.pos);
}
// adjoint: g_loc -> pose (pos, rot) and, optionally, the query point
template <class T> PLB_HD void inv_trans_bwd(V3<T> p, const Pose<T>& s, V3<T> gloc, PoseGrad<T>& gs, V3<T>* gp) {
    T n = plb_sqrt(s.rot.w * s.rot.w + s.rot.x * s.rot.x + s.rot.y * s.rot.y + s.rot.z * s.rot.z);
    Q4<T> qi; qi.w = s.rot.w / n; qi.x = -s.rot.x / n; qi.y = -s.rot.y / n; qi.z = -s.rot.z / n;
    V3<T> d = p - s.pos;
    V3<T> gd = qrot_bwd_v(qi, gloc);
    gs.pos -= gd;
    if (gp) *gp += gd;
    Q4<T> gqi; gqi.w = gqi.x = gqi.y = gqi.z = T(0);
    qrot_bwd_q(qi, d, gloc, gqi);
    // qi = c / |c|, c = conj(rot):  gc = (gqi - qi (qi . gqi)) / n
    T dq = qi.w * gqi.w + qi.x * gqi.x + qi.y * gqi.y + qi.z * gqi.z;
    gs.rot.w += (gqi.w - qi.w * dq) / n;
    gs.rot.x -= (gqi.x - qi.x * dq) / n;
    gs.rot.y -= (gqi.y - qi.y * dq) / n;
    gs.rot.z -= (gqi.z - qi.z * dq) / n;
}

// ---------------------------------------------------------------- small helpers (eps 1e-14 `length`)
template <class T> PLB_HD T plen3(V3<T> v) { return plb_sqrt(dot(v, v) + T(1e-14)); }
template <class T> PLB_HD T plen2(T a, T b) { return plb_sqrt(a * a + b * b + T(1e-14)); }
// adjoint of n = v / sqrt(v.v + eps): gv = gn / L - v (v . gn) / L^3
template <class T> PLB_HD V3<T> normalize3_vjp(V3<T> v, T L, V3<T> gn) {
    T s = dot(v, gn) / (L * L * L);
    return (T(1) / L) * gn - s * v;
}

// ---------------------------------------------------------------- local (object-frame) shapes
// capsule along y: returns p2 (the point relative to the clamped axis) and dy2/dy0
template <class T> PLB_HD V3<T> capsule_p2(T h, V3<T> q, T& dy) {
    T y0 = q.y + h / T(2);
    T z = tmax(y0, T(0));
    T cl = tmin(z, h);
    T dcl = ((T(0) < y0) && (z < h)) ? T(1) : T(0);
    dy = T(1) - dcl;
    return mk3<T>(q.x, y0 - cl, q.z);
}

template <class T> PLB_HD_NOINLINE T local_sdf(PrimStatic<T> ps, T gap, V3<T> q) {
    switch (ps.type) {
    case PRIM_CAPSULE: case PRIM_ROLLINGPIN: {
        T dy; V3<T> p2 = capsule_p2(ps.p[0], q, dy);
        return plen3(p2) - ps.p[1];
    }
    case PRIM_CHOPSTICKS: {
        V3<T> p = mk3<T>(q.x, q.y + ps.p[0] / T(2), q.z);
        T dy;
        T a = plen3(capsule_p2(ps.p[0], mk3<T>(p.x - gap / T(2), p.y, p.z), dy)) - ps.p[1];
        T b = plen3(capsule_p2(ps.p[0], mk3<T>(p.x + gap / T(2), p.y, p.z), dy)) - ps.p[1];
        return tmin(a, b);
    }
    case PRIM_CYLINDER: {
        T l = plen2(q.x, q.z);
        T d0 = plb_abs(l) - ps.p[0], d1 = plb_abs(q.y) - ps.p[1];
        return tmin(tmax(d0, d1), T(0)) + plen2(tmax(d0, T(0)), tmax(d1, T(0)));
    }
    case PRIM_TORUS: {
        T l = plen2(q.x, q.z);
        return plen2(l - ps.p[0], q.y) - ps.p[1];
    }
    case PRIM_BOX: {
        T d0 = plb_abs(q.x) - ps.p[0], d1 = plb_abs(q.y) - ps.p[1], d2 = plb_abs(q.z) - ps.p[2];
        T out = plen3(mk3<T>(tmax(d0, T(0)), tmax(d1, T(0)), tmax(d2, T(0))));
        return out + tmin(tmax(d0, tmax(d1, d2)), T(0));
    }
    default: return T(0);
    }
}

template <class T> PLB_HD T sgn0(T x) { return x > T(0) ? T(1) : (x < T(0) ? T(-1) : T(0)); }

// gradient of local_sdf wrt q, scaled by g; also d/d gap (chopsticks)
template <class T> PLB_HD_NOINLINE V3<T> local_sdf_vjp(PrimStatic<T> ps, T gap, V3<T> q, T g, T& ggap) {
    switch (ps.type) {
    case PRIM_CAPSULE: case PRIM_ROLLINGPIN: {
        T dy; V3<T> p2 = capsule_p2(ps.p[0], q, dy);
        T L = plen3(p2);
        return mk3<T>(g * p2.x / L, g * p2.y / L * dy, g * p2.z / L);
    }
    case PRIM_CHOPSTICKS: {
        V3<T> p = mk3<T>(q.x, q.y + ps.p[0] / T(2), q.z);
        T dya, dyb;
        V3<T> pa = capsule_p2(ps.p[0], mk3<T>(p.x - gap / T(2), p.y, p.z), dya);
        V3<T> pb = capsule_p2(ps.p[0], mk3<T>(p.x + gap / T(2), p.y, p.z), dyb);
        T La = plen3(pa), Lb = plen3(pb);
        T a = La - ps.p[1], b = Lb - ps.p[1];
        if (a < b) { ggap += g * (-T(0.5)) * pa.x / La; return mk3<T>(g * pa.x / La, g * pa.y / La * dya, g * pa.z / La); }
        ggap += g * (T(0.5)) * pb.x / Lb;
        return mk3<T>(g * pb.x / Lb, g * pb.y / Lb * dyb, g * pb.z / Lb);
    }
    case PRIM_CYLINDER: {
        T l = plen2(q.x, q.z);
        T d0 = plb_abs(l) - ps.p[0], d1 = plb_abs(q.y) - ps.p[1];
        T A = tmax(d0, d1);
        T m0 = tmax(d0, T(0)), m1 = tmax(d1, T(0));
        T L2 = plen2(m0, m1);
        T gA = (A < T(0)) ? g : T(0);
        T gd0 = ((d1 < d0) ? gA : T(0)) + ((T(0) < d0) ? g * m0 / L2 : T(0));
        T gd1 = ((d1 < d0) ? T(0) : gA) + ((T(0) < d1) ? g * m1 / L2 : T(0));
        T gl = gd0 * sgn0(l);
        return mk3<T>(gl * q.x / l, gd1 * sgn0(q.y), gl * q.z / l);
    }
    case PRIM_TORUS: {
        T l = plen2(q.x, q.z);
        T a = l - ps.p[0];
        T L = plen2(a, q.y);
        T gl = g * a / L;
        return mk3<T>(gl * q.x / l, g * q.y / L, gl * q.z / l);
    }
    case PRIM_BOX: {
        T d[3] = {plb_abs(q.x) - ps.p[0], plb_abs(q.y) - ps.p[1], plb_abs(q.z) - ps.p[2]};
        T m[3] = {tmax(d[0], T(0)), tmax(d[1], T(0)), tmax(d[2], T(0))};
        T L = plen3(mk3<T>(m[0], m[1], m[2]));
        T in12 = tmax(d[1], d[2]);
        T A = tmax(d[0], in12);
        T gA = (A < T(0)) ? g : T(0);
        T gd[3];
        gd[0] = (in12 < d[0]) ? gA : T(0);
        T gin = (in12 < d[0]) ? T(0) : gA;
        gd[1] = (d[2] < d[1]) ? gin : T(0);
        gd[2] = (d[2] < d[1]) ? T(0) : gin;
        V3<T> out;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            T gi = gd[i] + ((T(0) < d[i]) ? g * m[i] / L : T(0));
            out[i] = gi * sgn0(q[i]);
        }
        return out;
    }
    default: return zero3<T>();
    }
}

template <class T> PLB_HD_NOINLINE V3<T> local_normal(PrimStatic<T> ps, T gap, V3<T> q) {
    switch (ps.type) {
    case PRIM_CAPSULE: case PRIM_ROLLINGPIN: {
        T dy; V3<T> p2 = capsule_p2(ps.p[0], q, dy);
        return (T(1) / plen3(p2)) * p2;
    }
    case PRIM_CHOPSTICKS: {
        V3<T> p = mk3<T>(q.x, q.y + ps.p[0] / T(2), q.z);
        T dy;
        V3<T> pa = capsule_p2(ps.p[0], mk3<T>(p.x - gap / T(2), p.y, p.z), dy);
        V3<T> pb = capsule_p2(ps.p[0], mk3<T>(p.x + gap / T(2), p.y, p.z), dy);
        T La = plen3(pa), Lb = plen3(pb);
        return (La - ps.p[1] <= Lb - ps.p[1]) ? (T(1) / La) * pa : (T(1) / Lb) * pb;
    }
    case PRIM_CYLINDER: {
        T l = plen2(q.x, q.z);
        T d0 = l - ps.p[0], d1 = plb_abs(q.y) - ps.p[1];
        T f = (d0 > d1) ? T(1) : T(0);
        T inside = (((d0 > d1) ? d0 : d1) <= T(0)) ? T(1) : T(0);
        T n20 = tmax(d0, T(0)) + inside * f, n21 = tmax(d1, T(0)) + inside * (T(1) - f);
        T L2 = plen2(n20, n21);
        T a0 = n20 / L2, a1 = n21 / L2;
        T sg = (q.y >= T(0)) ? T(1) : T(-1);
        V3<T> n3 = mk3<T>(q.x / l * a0, a1 * sg, q.z / l * a0);
        return (T(1) / plen3(n3)) * n3;
    }
    case PRIM_TORUS: {
        T l = plen2(q.x, q.z);
        T a = l - ps.p[0];
        T L = plen2(a, q.y);
        V3<T> n3 = mk3<T>(q.x / l * (a / L), q.y / L, q.z / l * (a / L));
        return (T(1) / plen3(n3)) * n3;
    }
    case PRIM_BOX: {
        const T d = T(1e-4);
        V3<T> n;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            V3<T> inc = q, dec = q;
            inc[i] += d; dec[i] -= d;
            n[i] = (T(0.5) / d) * (local_sdf(ps, gap, inc) - local_sdf(ps, gap, dec));
        }
        return (T(1) / plen3(n)) * n;
    }
    default: return mk3<T>(T(0), T(1), T(0));
    }
}

// adjoint of local_normal wrt q (and gap)
template <class T> PLB_HD_NOINLINE V3<T> local_normal_vjp(PrimStatic<T> ps, T gap, V3<T> q, V3<T> gn, T& ggap) {
    switch (ps.type) {
    case PRIM_CAPSULE: case PRIM_ROLLINGPIN: {
        T dy; V3<T> p2 = capsule_p2(ps.p[0], q, dy);
        V3<T> g2 = normalize3_vjp(p2, plen3(p2), gn);
        return mk3<T>(g2.x, g2.y * dy, g2.z);
    }
    case PRIM_CHOPSTICKS: {
        V3<T> p = mk3<T>(q.x, q.y + ps.p[0] / T(2), q.z);
        T dya, dyb;
        V3<T> pa = capsule_p2(ps.p[0], mk3<T>(p.x - gap / T(2), p.y, p.z), dya);
        V3<T> pb = capsule_p2(ps.p[0], mk3<T>(p.x + gap / T(2), p.y, p.z), dyb);
        T La = plen3(pa), Lb = plen3(pb);
        if (La - ps.p[1] <= Lb - ps.p[1]) {
            V3<T> g2 = normalize3_vjp(pa, La, gn);
            ggap += -T(0.5) * g2.x;
            return mk3<T>(g2.x, g2.y * dya, g2.z);
        }
        V3<T> g2 = normalize3_vjp(pb, Lb, gn);
        ggap += T(0.5) * g2.x;
        return mk3<T>(g2.x, g2.y * dyb, g2.z);
    }
    case PRIM_CYLINDER: {
        T l = plen2(q.x, q.z);
        T d0 = l - ps.p[0], d1 = plb_abs(q.y) - ps.p[1];
        T f = (d0 > d1) ? T(1) : T(0);
        T inside = (((d0 > d1) ? d0 : d1) <= T(0)) ? T(1) : T(0);
        T n20 = tmax(d0, T(0)) + inside * f, n21 = tmax(d1, T(0)) + inside * (T(1) - f);
        T L2 = plen2(n20, n21);
        T a0 = n20 / L2, a1 = n21 / L2;
        T sg = (q.y >= T(0)) ? T(1) : T(-1);
        T p2x = q.x / l, p2z = q.z / l;
        V3<T> n3 = mk3<T>(p2x * a0, a1 * sg, p2z * a0);
        V3<T> g3 = normalize3_vjp(n3, plen3(n3), gn);
        T gp2x = g3.x * a0, gp2z = g3.z * a0;
        T ga0 = g3.x * p2x + g3.z * p2z, ga1 = g3.y * sg;
        // (a0, a1) = (n20, n21) / L2
        T s = (n20 * ga0 + n21 * ga1) / (L2 * L2 * L2);
        T gn20 = ga0 / L2 - s * n20, gn21 = ga1 / L2 - s * n21;
        T gd0 = (T(0) < d0) ? gn20 : T(0);
        T gd1 = (T(0) < d1) ? gn21 : T(0);
        // p2 = p / l ; l = plen2(p)
        T gl = gd0 - (gp2x * q.x + gp2z * q.z) / (l * l);
        return mk3<T>(gp2x / l + gl * q.x / l, gd1 * sgn0(q.y), gp2z / l + gl * q.z / l);
    }
    case PRIM_TORUS: {
        T l = plen2(q.x, q.z);
        T a = l - ps.p[0];
        T L = plen2(a, q.y);
        T n20 = a / L, n21 = q.y / L;
        T x2x = q.x / l, x2z = q.z / l;
        V3<T> n3 = mk3<T>(x2x * n20, n21, x2z * n20);
        V3<T> g3 = normalize3_vjp(n3, plen3(n3), gn);
        T gx2x = g3.x * n20, gx2z = g3.z * n20;
        T gn20 = g3.x * x2x + g3.z * x2z, gn21 = g3.y;
        T s = (a * gn20 + q.y * gn21) / (L * L * L);
        T ga = gn20 / L - s * a, gqy = gn21 / L - s * q.y;
        T gl = ga - (gx2x * q.x + gx2z * q.z) / (l * l);
        return mk3<T>(gx2x / l + gl * q.x / l, gqy, gx2z / l + gl * q.z / l);
    }
    case PRIM_BOX: {
        const T d = T(1e-4);
        V3<T> n;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            V3<T> inc = q, dec = q;
            inc[i] += d; dec[i] -= d;
            n[i] = (T(0.5) / d) * (local_sdf(ps, gap, inc) - local_sdf(ps, gap, dec));
        }
        V3<T> graw = normalize3_vjp(n, plen3(n), gn);
        V3<T> out = zero3<T>();
        T dummy = T(0);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            V3<T> inc = q, dec = q;
            inc[i] += d; dec[i] -= d;
            T gi = (T(0.5) / d) * graw[i];
            out += local_sdf_vjp(ps, gap, inc, gi, dummy);
            out -= local_sdf_vjp(ps, gap, dec, gi, dummy);
        }
        return out;
    }
    default: return zero3<T>();
    }
}

// ---------------------------------------------------------------- world-frame sdf (used by the loss too)
template <class T> PLB_HD T prim_sdf(const PrimStatic<T>& ps, const Pose<T>& s, V3<T> p) {
    if (ps.type == PRIM_SPHERE) return plen3(p - s.pos) - ps.p[0];
    return local_sdf(ps, s.gap, inv_trans(p, s));
}
// adjoint: g (scalar) -> pose, and optionally the query point
template <class T> PLB_HD void prim_sdf_vjp(const PrimStatic<T>& ps, const Pose<T>& s, V3<T> p, T g, PoseGrad<T>& gs, V3<T>* gp) {
    if (ps.type == PRIM_SPHERE) {
        V3<T> d = p - s.pos;
        V3<T> gd = (g / plen3(d)) * d;
        gs.pos -= gd;
        if (gp) *gp += gd;
        return;
    }
    V3<T> loc = inv_trans(p, s);
    V3<T> gloc = local_sdf_vjp(ps, s.gap, loc, g, gs.gap);
    inv_trans_bwd(p, s, gloc, gs, gp);
}

// ---------------------------------------------------------------- collide (primive_base.py:91-115)
// Forward.  Returns the new grid velocity; `taken` tells whether the contact branch ran.
template <class T>
PLB_HD V3<T> prim_collide(const PrimStatic<T>& ps, const Pose<T>& s0, const Pose<T>& s1, V3<T> gpos, V3<T> v, T dt, bool& taken) {
    V3<T> loc = inv_trans(gpos, s0);
    T dist = (ps.type == PRIM_SPHERE) ? plen3(gpos - s0.pos) - ps.p[0] : local_sdf(ps, s0.gap, loc);
    T e = plb_exp(-dist * ps.softness);
    T infl = tmin(e, T(1));
    taken = (ps.softness > T(0) && infl > T(0.1)) || dist <= T(0);
    if (!taken) return v;
    V3<T> D;
    if (ps.type == PRIM_SPHERE) { V3<T> d = gpos - s0.pos; D = (T(1) / plen3(d)) * d; }
    else D = qrot(s0.rot, local_normal(ps, s0.gap, loc));
    V3<T> cv = (T(1) / dt) * (qrot(s1.rot, loc) + s1.pos - gpos);
    V3<T> iv = v - cv;
    T nc = dot(iv, D);
    T t = tmin(nc, T(0));
    V3<T> vt = iv - t * D;
    T vt2 = dot(vt, vt);
    T vtn = plb_sqrt(vt2 + T(1e-8));
    T fr = tmax(T(0), vtn + nc * ps.friction);
    bool flag = (nc < T(0)) && (plb_sqrt(vt2) > T(1e-30));
    V3<T> vtf = flag ? (fr / vtn) * vt : vt;
    return cv + (T(1) - infl) * iv + infl * vtf;
}

// Adjoint.  Given v (the input velocity of this collide) and gout (adjoint of its output) returns the adjoint
// of v and accumulates pose adjoints for frame f (g0) and f+1 (g1).
template <class T>
PLB_HD V3<T> prim_collide_bwd(const PrimStatic<T>& ps, const Pose<T>& s0, const Pose<T>& s1, V3<T> gpos, V3<T> v, T dt,
                              V3<T> gout, PoseGrad<T>& g0, PoseGrad<T>& g1, bool& taken) {
    V3<T> loc = inv_trans(gpos, s0);
    const bool sphere = (ps.type == PRIM_SPHERE);
    V3<T> dsp = gpos - s0.pos;
    T Lsp = plen3(dsp);
    T dist = sphere ? Lsp - ps.p[0] : local_sdf(ps, s0.gap, loc);
    T e = plb_exp(-dist * ps.softness);
    T infl = tmin(e, T(1));
    taken = (ps.softness > T(0) && infl > T(0.1)) || dist <= T(0);
    if (!taken) return gout;
    V3<T> nl = zero3<T>();
    V3<T> D;
    if (sphere) D = (T(1) / Lsp) * dsp;
    else { nl = local_normal(ps, s0.gap, loc); D = qrot(s0.rot, nl); }
    V3<T> cv = (T(1) / dt) * (qrot(s1.rot, loc) + s1.pos - gpos);
    V3<T> iv = v - cv;
    T nc = dot(iv, D);
    T t = tmin(nc, T(0));
    V3<T> vt = iv - t * D;
    T vt2 = dot(vt, vt);
    T vtn = plb_sqrt(vt2 + T(1e-8));
    T y = vtn + nc * ps.friction;
    T fr = tmax(T(0), y);
    bool flag = (nc < T(0)) && (plb_sqrt(vt2) > T(1e-30));
    V3<T> vtsel = flag ? (fr / vtn) * vt : vt;
    // ---- backward
    V3<T> gcv = gout;
    V3<T> giv = (T(1) - infl) * gout;
    T ginfl = dot(gout, vtsel) - dot(gout, iv);
    V3<T> gsel = infl * gout;
    V3<T> gvt;
    T gnc = T(0);
    if (flag) {
        T ratio = fr / vtn;
        gvt = ratio * gsel;
        T gratio = dot(gsel, vt);
        T gfr = gratio / vtn;
        T gvtn = -gratio * fr / (vtn * vtn);
        T gy = (y < T(0)) ? T(0) : gfr;              // tmax(0, y): to y unless y < 0
        gvtn += gy;
        gnc += ps.friction * gy;
        gvt += (gvtn / vtn) * vt;
    } else {
        gvt = gsel;
    }
    giv += gvt;
    T gt = -dot(gvt, D);
    V3<T> gD = (-t) * gvt;
    if (nc < T(0)) gnc += gt;                         // tmin(nc, 0): to nc iff nc < 0
    giv += gnc * D;
    gD += gnc * iv;
    V3<T> gv = giv;
    gcv -= giv;
    T ge = (e < T(1)) ? ginfl : T(0);                 // tmin(e, 1): to e iff e < 1
    T gdist = -ps.softness * e * ge;
    // cv = (qrot(rot1, loc) + pos1 - gpos) / dt
    V3<T> gc = (T(1) / dt) * gcv;
    g1.pos += gc;
    qrot_bwd_q(s1.rot, loc, gc, g1.rot);
    V3<T> gloc = qrot_bwd_v(s1.rot, gc);
    if (sphere) {
        // dist = |d| - R ; D = d / |d|
        V3<T> gd = (gdist / Lsp) * dsp + normalize3_vjp(dsp, Lsp, gD);
        g0.pos -= gd;
    } else {
        qrot_bwd_q(s0.rot, nl, gD, g0.rot);
        V3<T> gnl = qrot_bwd_v(s0.rot, gD);
        gloc += local_normal_vjp(ps, s0.gap, loc, gnl, g0.gap);
        gloc += local_sdf_vjp(ps, s0.gap, loc, gdist, g0.gap);
    }
    inv_trans_bwd(gpos, s0, gloc, g0, (V3<T>*)nullptr);
    return gv;
}

}  // namespace plb
