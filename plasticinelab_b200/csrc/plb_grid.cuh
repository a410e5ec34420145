// Per-node grid operator and its adjoint (plb/engine/mpm_simulator.py:189-221).
//
//   m > 1e-12:  v = v_in / m;  v += dt * gravity * 30;  v = collide_k(v) for every primitive in order;
//               box boundary with bound = 3, axis by axis ON THE MUTATED v (ground friction modes for axis 1).
// Nodes with m <= 1e-12 output 0 and receive no gradient.
#pragma once
#include "plb_particle.cuh"
#include "plb_primitives.cuh"

namespace plb {

template <class T> struct PrimSet {
    PrimStatic<T> s[PLB_MAX_PRIM];
};

// ---- the boundary part, forward, recording what the adjoint needs
template <class T> struct BoundaryTape {
    unsigned mask;        // bit d: low-side rule fired on axis d; bit 4+d: high-side rule fired
    V3<T> v_fric;         // v before the friction rule (axis 1, 0 < ground_friction < 10)
};

template <class T>
PLB_HD V3<T> boundary_forward(const SimConst<T>& P, int ix, int iy, int iz, V3<T> v, BoundaryTape<T>* tape) {
    const int bound = 3;
    const int I[3] = {ix, iy, iz};
    const T gf = P.ground_friction;
    unsigned mask = 0;
    V3<T> vf = zero3<T>();
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (I[d] < bound && v[d] < T(0)) {
            mask |= 1u << d;
            if (d != 1 || gf == T(0)) {
                v[d] = T(0);
            } else if (gf < T(10)) {
                vf = v;
                V3<T> tiny = mk3<T>(T(ix) * T(1e-30), T(iy) * T(1e-30), T(iz) * T(1e-30));
                T lin = v.y + T(1e-30);
                V3<T> vit = mk3<T>(v.x - tiny.x, v.y - lin - tiny.y, v.z - tiny.z);
                T lit = plb_sqrt(dot(vit, vit) + T(1e-8));
                T s = tmax(T(1) + gf * lin / lit, T(0));
                v = s * (vit + tiny);
                v.y = T(0);
            } else {
                v = zero3<T>();
            }
        }
        if (I[d] > P.n_grid - bound && v[d] > T(0)) { mask |= 1u << (4 + d); v[d] = T(0); }
    }
    if (tape) { tape->mask = mask; tape->v_fric = vf; }
    return v;
}

template <class T>
PLB_HD V3<T> boundary_backward(const SimConst<T>& P, int ix, int iy, int iz, const BoundaryTape<T>& tape, V3<T> g) {
    const T gf = P.ground_friction;
#pragma unroll
    for (int d = 2; d >= 0; d--) {
        if (tape.mask & (1u << (4 + d))) g[d] = T(0);
        if (tape.mask & (1u << d)) {
            if (d != 1 || gf == T(0)) {
                g[d] = T(0);
            } else if (gf < T(10)) {
                V3<T> v = tape.v_fric;
                V3<T> tiny = mk3<T>(T(ix) * T(1e-30), T(iy) * T(1e-30), T(iz) * T(1e-30));
                T lin = v.y + T(1e-30);
                V3<T> vit = mk3<T>(v.x - tiny.x, v.y - lin - tiny.y, v.z - tiny.z);
                T lit = plb_sqrt(dot(vit, vit) + T(1e-8));
                T a = T(1) + gf * lin / lit;
                T s = tmax(a, T(0));
                V3<T> gnew = mk3<T>(g.x, T(0), g.z);             // v[1] is overwritten with 0
                T gs = dot(gnew, vit + tiny);
                V3<T> gvit = s * gnew;
                T ga = (T(0) < a) ? gs : T(0);                    // tmax(a, 0): to a iff 0 < a
                T glin = gf * ga / lit;
                T glit = -gf * ga * lin / (lit * lit);
                gvit += (glit / lit) * vit;
                // vit = v - lin * e1 - tiny ; lin = v.y + 1e-30
                glin -= gvit.y;
                g = gvit;
                g.y += glin;
            } else {
                g = zero3<T>();
            }
        }
    }
    return g;
}

// Forward of one node.  in4 = (momentum xyz, mass).  poses: [n_prim] for frame f and f+1.
template <class T>
PLB_HD V3<T> grid_node_forward(const SimConst<T>& P, const PrimSet<T>& prims, const Pose<T>* s0, const Pose<T>* s1,
                               int ix, int iy, int iz, Vec4<T> in4) {
    if (!(in4.w > T(1e-12))) return zero3<T>();
    T inv_m = T(1) / in4.w;
    V3<T> v = mk3<T>(inv_m * in4.x + P.grav_dv[0], inv_m * in4.y + P.grav_dv[1], inv_m * in4.z + P.grav_dv[2]);
    V3<T> gpos = mk3<T>(T(ix) * P.dx, T(iy) * P.dx, T(iz) * P.dx);
    for (int k = 0; k < P.n_prim; k++) {
        bool taken;
        v = prim_collide(prims.s[k], s0[k], s1[k], gpos, v, P.dt, taken);
    }
    return boundary_forward<T>(P, ix, iy, iz, v, nullptr);
}

// Adjoint of one node: gout = adjoint of v_out.  Returns the adjoint of in4 and accumulates pose adjoints
// into g0[k], g1[k] (caller-zeroed, one pair per primitive); `touched` has bit k set when primitive k's
// contact branch ran on this node.
template <class T>
PLB_HD Vec4<T> grid_node_backward(const SimConst<T>& P, const PrimSet<T>& prims, const Pose<T>* s0, const Pose<T>* s1,
                                  int ix, int iy, int iz, Vec4<T> in4, V3<T> gout,
                                  PoseGrad<T>* g0, PoseGrad<T>* g1, unsigned& touched) {
    touched = 0;
    if (!(in4.w > T(1e-12))) return mk4<T>(T(0), T(0), T(0), T(0));
    T inv_m = T(1) / in4.w;
    V3<T> vin = mk3<T>(in4.x, in4.y, in4.z);
    V3<T> v = mk3<T>(inv_m * in4.x + P.grav_dv[0], inv_m * in4.y + P.grav_dv[1], inv_m * in4.z + P.grav_dv[2]);
    V3<T> gpos = mk3<T>(T(ix) * P.dx, T(iy) * P.dx, T(iz) * P.dx);
    V3<T> vstack[PLB_MAX_PRIM];
    for (int k = 0; k < P.n_prim; k++) {
        vstack[k] = v;
        bool taken;
        v = prim_collide(prims.s[k], s0[k], s1[k], gpos, v, P.dt, taken);
    }
    BoundaryTape<T> tape;
    boundary_forward<T>(P, ix, iy, iz, v, &tape);
    V3<T> g = boundary_backward<T>(P, ix, iy, iz, tape, gout);
    for (int k = P.n_prim - 1; k >= 0; k--) {
        bool taken;
        g = prim_collide_bwd(prims.s[k], s0[k], s1[k], gpos, vstack[k], P.dt, g, g0[k], g1[k], taken);
        if (taken) touched |= 1u << k;
    }
    // v = inv_m * v_in + const ; inv_m = 1 / m
    T ginv = dot(g, vin);
    return mk4<T>(inv_m * g.x, inv_m * g.y, inv_m * g.z, -ginv * inv_m * inv_m);
}

}  // namespace plb
