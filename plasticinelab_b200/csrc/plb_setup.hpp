// plb_config / plb_primitive_desc -> kernel constants.  Shared by the engine (plb_engine.cu) and the host emulation
// used by the CPU tests (tests/host/emul.cpp) so both see identical derived constants.
#pragma once
#include <vector>
#include "../../include/plb_b200.h"
#include "plb_bodies.cuh"
#include "plb_kinematics.hpp"

namespace plb {

template <class T> inline SimConst<T> make_simconst(const plb_config& c) {
    SimConst<T> P{};
    P.dx = (T)c.dx; P.inv_dx = (T)(1.0 / c.dx); P.dt = (T)c.dt; P.p_vol = (T)c.p_vol; P.p_mass = (T)c.p_mass;
    P.stress_scale = (T)(-c.dt * c.p_vol * 4.0 * (1.0 / c.dx) * (1.0 / c.dx));
    P.x_hi = (T)(1.0 - 3.0 * c.dx);
    for (int d = 0; d < 3; d++) P.grav_dv[d] = (T)(c.dt * c.gravity[d] * 30.0);
    P.ground_friction = (T)c.ground_friction;
    P.n_grid = c.n_grid; P.n_particles = c.n_particles; P.n_prim = c.n_primitives;
    double mu = c.E / (2 * (1 + c.nu)), lam = c.E * c.nu / ((1 + c.nu) * (1 - 2 * c.nu));
    P.mu = (T)mu; P.lam = (T)lam; P.yield_stress = (T)c.yield_stress;
    return P;
}

template <class T> inline PrimStatic<T> make_primstatic(const plb_primitive_desc& d, double softness) {
    PrimStatic<T> s{};
    s.type = d.type; s.movable = d.action_dim > 0;
    for (int i = 0; i < 4; i++) s.p[i] = (T)d.params[i];
    s.friction = (T)d.friction; s.softness = (T)softness;
    return s;
}

inline kin::Desc make_kindesc(const plb_primitive_desc& d) {
    kin::Desc kd{};
    kd.type = d.type; kd.action_dim = d.action_dim; kd.minimal_gap = d.minimal_gap;
    for (int i = 0; i < 3; i++) { kd.lower[i] = d.lower_bound[i]; kd.upper[i] = d.upper_bound[i]; }
    for (int i = 0; i < PLB_MAX_ACTION_DIM; i++) kd.action_scale[i] = d.action_scale[i];
    return kd;
}

}  // namespace plb
