/* Plain-C float64 port of the reference substep and its adjoint -- ORACLE / CPU BASELINE, TEST INFRASTRUCTURE ONLY.
 *
 * Restates, kernel by kernel and with dense grid sweeps like the reference, what plb/engine/mpm_simulator.py computes:
 *   clear_grid :60-70, compute_F_tmp :82-85, svd :87-90, compute_von_mises :124-141, p2g :157-184, grid_op :189-221
 *   (with Sphere.collide, plb/engine/primitive/primive_base.py:91-115 + primitives.py:17-34), g2p :223-242, and the
 *   adjoint order of substep_grad :260-278 (g2p.grad, grid_op.grad, p2g.grad, svd_grad :92-115, compute_F_tmp.grad).
 * Parallelism: OpenMP over particles / nodes with atomic adds on the grid, which is what Taichi's x64 backend does for
 * `+=` on fields.  Only Sphere primitives are ported (Move / TripleMove / Rope-style scenes and both bench workloads);
 * the torch oracle (plb_oracle.py) remains the checker for every other shape, and this file is validated against it in
 * tests/test_oracle.py.  Used by bench.py as the CPU baseline (`cpu_baseline.kind = "port"`) because the reference's
 * own runtime (Taichi 0.7.14) is not installable here.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n, n_grid, n_prim;
    double dx, inv_dx, dt, p_vol, p_mass, mu, lam, yield_stress, ground_friction;
    double grav_dv[3];                  /* dt * gravity * 30 */
    double radius[8], friction[8], softness;
} oc_params;

/* ------------------------------------------------------------------ 3x3 helpers (row major double[9]) */
static void mm(const double* a, const double* b, double* r) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i*3+j] = a[i*3]*b[j] + a[i*3+1]*b[3+j] + a[i*3+2]*b[6+j]; }
static void mmT(const double* a, const double* b, double* r) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i*3+j] = a[i*3]*b[j*3] + a[i*3+1]*b[j*3+1] + a[i*3+2]*b[j*3+2]; }
static void mTm(const double* a, const double* b, double* r) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i*3+j] = a[i]*b[j] + a[3+i]*b[3+j] + a[6+i]*b[6+j]; }
static double det3(const double* a) { return a[0]*(a[4]*a[8]-a[5]*a[7]) - a[1]*(a[3]*a[8]-a[5]*a[6]) + a[2]*(a[3]*a[7]-a[4]*a[6]); }
static void cof3(const double* a, double* c) {
    c[0]=a[4]*a[8]-a[5]*a[7]; c[1]=a[5]*a[6]-a[3]*a[8]; c[2]=a[3]*a[7]-a[4]*a[6];
    c[3]=a[2]*a[7]-a[1]*a[8]; c[4]=a[0]*a[8]-a[2]*a[6]; c[5]=a[1]*a[6]-a[0]*a[7];
    c[6]=a[1]*a[5]-a[2]*a[4]; c[7]=a[2]*a[3]-a[0]*a[5]; c[8]=a[0]*a[4]-a[1]*a[3];
}

/* one-sided Jacobi SVD, convention of ti.svd: det U = det V = +1, descending, sign on the last singular value */
static void svd3(const double* F, double* U, double* sig, double* V) {
    double a[3][3], v[3][3];
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) { a[c][r] = F[r*3+c]; v[c][r] = (r == c); }
    for (int sweep = 0; sweep < 30; sweep++) {
        int rot = 0;
        for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
            double al = 0, be = 0, ga = 0;
            for (int r = 0; r < 3; r++) { al += a[p][r]*a[p][r]; be += a[q][r]*a[q][r]; ga += a[p][r]*a[q][r]; }
            if (ga*ga > 1e-31*al*be) {
                rot = 1;
                double zeta = (be - al) / (2*ga), t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta*zeta));
                double c = 1/sqrt(1 + t*t), s = c*t;
                for (int r = 0; r < 3; r++) {
                    double ap = c*a[p][r] - s*a[q][r], aq = s*a[p][r] + c*a[q][r]; a[p][r] = ap; a[q][r] = aq;
                    double vp = c*v[p][r] - s*v[q][r], vq = s*v[p][r] + c*v[q][r]; v[p][r] = vp; v[q][r] = vq;
                }
            }
        }
        if (!rot) break;
    }
    double n[3]; int o[3] = {0, 1, 2};
    for (int c = 0; c < 3; c++) n[c] = a[c][0]*a[c][0] + a[c][1]*a[c][1] + a[c][2]*a[c][2];
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2 - i; j++) if (n[o[j]] < n[o[j+1]]) { int t = o[j]; o[j] = o[j+1]; o[j+1] = t; }
    double A[3][3], W[3][3];
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) { A[c][r] = a[o[c]][r]; W[c][r] = v[o[c]][r]; }
    double dv = W[0][0]*(W[1][1]*W[2][2]-W[1][2]*W[2][1]) - W[0][1]*(W[1][0]*W[2][2]-W[1][2]*W[2][0]) + W[0][2]*(W[1][0]*W[2][1]-W[1][1]*W[2][0]);
    if (dv < 0) for (int r = 0; r < 3; r++) { W[2][r] = -W[2][r]; A[2][r] = -A[2][r]; }
    double s0 = sqrt(n[o[0]]), s1 = sqrt(n[o[1]]), u0[3], u1[3], u2[3];
    for (int r = 0; r < 3; r++) { u0[r] = A[0][r]/s0; u1[r] = A[1][r]/s1; }
    u2[0] = u0[1]*u1[2]-u0[2]*u1[1]; u2[1] = u0[2]*u1[0]-u0[0]*u1[2]; u2[2] = u0[0]*u1[1]-u0[1]*u1[0];
    sig[0] = s0; sig[1] = s1; sig[2] = u2[0]*A[2][0] + u2[1]*A[2][1] + u2[2]*A[2][2];
    for (int r = 0; r < 3; r++) { U[r*3] = u0[r]; U[r*3+1] = u1[r]; U[r*3+2] = u2[r]; V[r*3] = W[0][r]; V[r*3+1] = W[1][r]; V[r*3+2] = W[2][r]; }
}

static double clampf(double a) { return a >= 0 ? (a > 1e-6 ? a : 1e-6) : (a < -1e-6 ? a : -1e-6); }

/* backward_svd, mpm_simulator.py:97-115 (gsig: adjoint of the diagonal) */
static void svd3_bwd(const double* gU, const double* gs, const double* gV, const double* U, const double* sig, const double* V, double* gF) {
    double a[9], b[9], in[9], t[9];
    mTm(U, gU, a); mTm(V, gV, b);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        if (i == j) { in[i*3+j] = gs[i]; continue; }
        double Fm = 1.0 / clampf(sig[j]*sig[j] - sig[i]*sig[i]);
        in[i*3+j] = Fm*(a[i*3+j]-a[j*3+i])*sig[j] + sig[i]*Fm*(b[i*3+j]-b[j*3+i]);
    }
    mm(U, in, t); mmT(t, V, gF);
}

typedef struct { double Ft[9], U[9], V[9], nF[9], M[9], sig[3], eh[3], e[3], n, J, c; int yield; } p2g_keep;

static void p2g_particle(const oc_params* P, const double* C, const double* F, double* nF, double* aff, p2g_keep* k) {
    double A[9], Ft[9], U[9], V[9], s[3];
    for (int i = 0; i < 9; i++) A[i] = P->dt*C[i];
    A[0] += 1; A[4] += 1; A[8] += 1;
    mm(A, F, Ft);
    svd3(Ft, U, s, V);
    double eps[3], mean = 0, eh[3], e[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) { double sc = 0.05 < s[i] ? s[i] : 0.05; eps[i] = log(sc); mean += eps[i]/3; }
    double nn = 1e-8;
    for (int i = 0; i < 3; i++) { eh[i] = eps[i]-mean; nn += eh[i]*eh[i]; }
    nn = sqrt(nn);
    double c = P->yield_stress/(2*P->mu), dg = nn - c;
    int yield = dg > 0;
    if (yield) {
        double UE[9];
        for (int i = 0; i < 3; i++) e[i] = exp(eps[i] - dg/nn*eh[i]);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) UE[i*3+j] = U[i*3+j]*e[j];
        mmT(UE, V, nF);
    } else memcpy(nF, Ft, sizeof(Ft));
    double J = det3(nF), r[9], M[9], st[9];
    mmT(U, V, r);
    for (int i = 0; i < 9; i++) M[i] = nF[i]-r[i];
    mmT(M, nF, st);
    double scale = -P->dt*P->p_vol*4*P->inv_dx*P->inv_dx;
    for (int i = 0; i < 9; i++) aff[i] = scale*(2*P->mu*st[i] + ((i % 4 == 0) ? P->lam*J*(J-1) : 0)) + P->p_mass*C[i];
    if (k) { memcpy(k->Ft, Ft, 72); memcpy(k->U, U, 72); memcpy(k->V, V, 72); memcpy(k->nF, nF, 72); memcpy(k->M, M, 72);
             memcpy(k->sig, s, 24); memcpy(k->eh, eh, 24); memcpy(k->e, e, 24); k->n = nn; k->J = J; k->c = c; k->yield = yield; }
}

static void p2g_particle_bwd(const oc_params* P, const double* C, const double* F, const p2g_keep* k, const double* gaff, const double* gFn, double* gC, double* gF) {
    double S[9], gM[9], gNF[9], t[9], cf[9], gU[9], gV[9], gs[3] = {0, 0, 0}, gFt[9];
    double scale = -P->dt*P->p_vol*4*P->inv_dx*P->inv_dx;
    for (int i = 0; i < 9; i++) { gC[i] = P->p_mass*gaff[i]; S[i] = scale*gaff[i]; }
    mm(S, k->nF, gM);
    mTm(S, k->M, t);
    for (int i = 0; i < 9; i++) { gM[i] *= 2*P->mu; gNF[i] = gM[i] + 2*P->mu*t[i] + gFn[i]; }
    double gJ = P->lam*(2*k->J-1)*(S[0]+S[4]+S[8]);
    cof3(k->nF, cf);
    for (int i = 0; i < 9; i++) gNF[i] += gJ*cf[i];
    mm(gM, k->V, gU); mTm(gM, k->U, gV);
    for (int i = 0; i < 9; i++) { gU[i] = -gU[i]; gV[i] = -gV[i]; }
    if (k->yield) {
        double VE[9], UE[9], a[9], b[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { VE[i*3+j] = k->V[i*3+j]*k->e[j]; UE[i*3+j] = k->U[i*3+j]*k->e[j]; }
        mm(gNF, VE, a); mTm(gNF, UE, b);
        for (int i = 0; i < 9; i++) { gU[i] += a[i]; gV[i] += b[i]; }
        mTm(k->U, gNF, a); mm(a, k->V, b);
        double ge1[3] = {b[0]*k->e[0], b[4]*k->e[1], b[8]*k->e[2]}, ge[3], geh[3], gk = 0, n = k->n, kk = 1 - k->c/n;
        for (int i = 0; i < 3; i++) { ge[i] = ge1[i]; geh[i] = -kk*ge1[i]; gk -= ge1[i]*k->eh[i]; }
        double gn = gk*k->c/(n*n), gm = 0;
        for (int i = 0; i < 3; i++) { geh[i] += gn/n*k->eh[i]; gm += geh[i]/3; }
        for (int i = 0; i < 3; i++) { ge[i] += geh[i]-gm; gs[i] = (0.05 < k->sig[i]) ? ge[i]/k->sig[i] : 0; }
        memset(gFt, 0, sizeof(gFt));
    } else memcpy(gFt, gNF, sizeof(gFt));
    svd3_bwd(gU, gs, gV, k->U, k->sig, k->V, t);
    for (int i = 0; i < 9; i++) gFt[i] += t[i];
    mmT(gFt, F, t);
    for (int i = 0; i < 9; i++) gC[i] += P->dt*t[i];
    double A[9];
    for (int i = 0; i < 9; i++) A[i] = P->dt*C[i];
    A[0] += 1; A[4] += 1; A[8] += 1;
    mTm(A, gFt, gF);
}

static void stencil(const oc_params* P, const double* x, int* b, double* fx, double w[3][3]) {
    for (int d = 0; d < 3; d++) {
        double xs = x[d]*P->inv_dx; b[d] = (int)(xs - 0.5); double f = xs - b[d]; fx[d] = f;
        w[0][d] = 0.5*(1.5-f)*(1.5-f); w[1][d] = 0.75-(f-1)*(f-1); w[2][d] = 0.5*(f-0.5)*(f-0.5);
    }
}
static double dwt(int a, double f) { return a == 0 ? -(1.5-f) : (a == 1 ? -2*(f-1) : (f-0.5)); }

/* ------------------------------------------------------------------ grid operator (Sphere primitives) */
static void sphere_collide(const oc_params* P, int k, const double* s0, const double* s1, const double* g, double* v, int bwd, double* gout,
                           double* g0, double* g1) {
    /* forward: v <- collide(v); backward (bwd=1): v holds the INPUT velocity, gout the output adjoint (overwritten by the
       input adjoint); g0/g1 accumulate pose adjoints.  Sphere: sdf/normal in world frame (primitives.py:22-28). */
    double d[3] = {g[0]-s0[0], g[1]-s0[1], g[2]-s0[2]}, L = sqrt(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]+1e-14), dist = L - P->radius[k];
    double e = exp(-dist*P->softness), infl = e < 1 ? e : 1;
    if (!((P->softness > 0 && infl > 0.1) || dist <= 0)) return;
    double D[3] = {d[0]/L, d[1]/L, d[2]/L};
    /* collider_v: rel = qrot(conj(q0)/|q0|, g - p0); new = qrot(q1, rel) + p1 */
    double qn = sqrt(s0[3]*s0[3]+s0[4]*s0[4]+s0[5]*s0[5]+s0[6]*s0[6]);
    double qi[4] = {s0[3]/qn, -s0[4]/qn, -s0[5]/qn, -s0[6]/qn}, rel[3], np[3];
    #define QROT(q, vv, out) { double ux=(q)[1],uy=(q)[2],uz=(q)[3]; double c0=uy*(vv)[2]-uz*(vv)[1], c1=uz*(vv)[0]-ux*(vv)[2], c2=ux*(vv)[1]-uy*(vv)[0]; \
        double e0=uy*c2-uz*c1, e1=uz*c0-ux*c2, e2=ux*c1-uy*c0; (out)[0]=(vv)[0]+2*((q)[0]*c0+e0); (out)[1]=(vv)[1]+2*((q)[0]*c1+e1); (out)[2]=(vv)[2]+2*((q)[0]*c2+e2); }
    QROT(qi, d, rel);
    QROT(s1 + 3, rel, np);
    double cv[3], iv[3], nc = 0, vt[3], vt2 = 0;
    for (int i = 0; i < 3; i++) { cv[i] = (np[i] + s1[i] - g[i])/P->dt; iv[i] = v[i]-cv[i]; nc += iv[i]*D[i]; }
    double t = nc < 0 ? nc : 0;
    for (int i = 0; i < 3; i++) { vt[i] = iv[i]-t*D[i]; vt2 += vt[i]*vt[i]; }
    double vtn = sqrt(vt2+1e-8), y = vtn + nc*P->friction[k], fr = y < 0 ? 0 : y;
    int flag = (nc < 0) && (sqrt(vt2) > 1e-30);
    double sel[3];
    for (int i = 0; i < 3; i++) sel[i] = flag ? fr/vtn*vt[i] : vt[i];
    if (!bwd) { for (int i = 0; i < 3; i++) v[i] = cv[i] + (1-infl)*iv[i] + infl*sel[i]; return; }
    /* ---- adjoint */
    double gcv[3], giv[3], gsel[3], gvt[3], ginfl = 0, gnc = 0, gD[3], gt = 0;
    for (int i = 0; i < 3; i++) { gcv[i] = gout[i]; giv[i] = (1-infl)*gout[i]; ginfl += gout[i]*(sel[i]-iv[i]); gsel[i] = infl*gout[i]; }
    if (flag) {
        double ratio = fr/vtn, gr = 0;
        for (int i = 0; i < 3; i++) { gvt[i] = ratio*gsel[i]; gr += gsel[i]*vt[i]; }
        double gfr = gr/vtn, gvtn = -gr*fr/(vtn*vtn), gy = y < 0 ? 0 : gfr;
        gvtn += gy; gnc += P->friction[k]*gy;
        for (int i = 0; i < 3; i++) gvt[i] += gvtn/vtn*vt[i];
    } else for (int i = 0; i < 3; i++) gvt[i] = gsel[i];
    for (int i = 0; i < 3; i++) { giv[i] += gvt[i]; gt -= gvt[i]*D[i]; gD[i] = -t*gvt[i]; }
    if (nc < 0) gnc += gt;
    for (int i = 0; i < 3; i++) { giv[i] += gnc*D[i]; gD[i] += gnc*iv[i]; }
    for (int i = 0; i < 3; i++) { gout[i] = giv[i]; gcv[i] -= giv[i]; }
    double ge = e < 1 ? ginfl : 0, gdist = -P->softness*e*ge;
    double gc[3] = {gcv[0]/P->dt, gcv[1]/P->dt, gcv[2]/P->dt};
    for (int i = 0; i < 3; i++) g1[i] += gc[i];
    /* qrot adjoints */
    #define QROT_BWD_V(q, gg, out) { double ux=(q)[1],uy=(q)[2],uz=(q)[3]; double c0=uy*(gg)[2]-uz*(gg)[1], c1=uz*(gg)[0]-ux*(gg)[2], c2=ux*(gg)[1]-uy*(gg)[0]; \
        double e0=uy*c2-uz*c1, e1=uz*c0-ux*c2, e2=ux*c1-uy*c0; (out)[0]=(gg)[0]+2*(e0-(q)[0]*c0); (out)[1]=(gg)[1]+2*(e1-(q)[0]*c1); (out)[2]=(gg)[2]+2*(e2-(q)[0]*c2); }
    #define QROT_BWD_Q(q, vv, gg, gq) { double ux=(q)[1],uy=(q)[2],uz=(q)[3]; double c0=uy*(vv)[2]-uz*(vv)[1], c1=uz*(vv)[0]-ux*(vv)[2], c2=ux*(vv)[1]-uy*(vv)[0]; \
        (gq)[0] += 2*((gg)[0]*c0+(gg)[1]*c1+(gg)[2]*c2); \
        double x0=(vv)[1]*(gg)[2]-(vv)[2]*(gg)[1], x1=(vv)[2]*(gg)[0]-(vv)[0]*(gg)[2], x2=(vv)[0]*(gg)[1]-(vv)[1]*(gg)[0]; \
        double uv=ux*(vv)[0]+uy*(vv)[1]+uz*(vv)[2], gu_=(gg)[0]*ux+(gg)[1]*uy+(gg)[2]*uz, gv_=(gg)[0]*(vv)[0]+(gg)[1]*(vv)[1]+(gg)[2]*(vv)[2]; \
        (gq)[1] += 2*((q)[0]*x0 + uv*(gg)[0] + gu_*(vv)[0] - 2*gv_*ux); (gq)[2] += 2*((q)[0]*x1 + uv*(gg)[1] + gu_*(vv)[1] - 2*gv_*uy); \
        (gq)[3] += 2*((q)[0]*x2 + uv*(gg)[2] + gu_*(vv)[2] - 2*gv_*uz); }
    double gq1[4] = {0, 0, 0, 0}, grel[3], gqi[4] = {0, 0, 0, 0}, gd[3];
    QROT_BWD_Q(s1 + 3, rel, gc, gq1);
    for (int i = 0; i < 4; i++) g1[3+i] += gq1[i];
    QROT_BWD_V(s1 + 3, gc, grel);
    QROT_BWD_V(qi, grel, gd);
    QROT_BWD_Q(qi, d, grel, gqi);
    double dq = qi[0]*gqi[0]+qi[1]*gqi[1]+qi[2]*gqi[2]+qi[3]*gqi[3];
    g0[3] += (gqi[0]-qi[0]*dq)/qn; g0[4] -= (gqi[1]-qi[1]*dq)/qn; g0[5] -= (gqi[2]-qi[2]*dq)/qn; g0[6] -= (gqi[3]-qi[3]*dq)/qn;
    /* dist = |d| - R, D = d/|d| */
    double dg = (d[0]*gD[0]+d[1]*gD[1]+d[2]*gD[2])/(L*L*L);
    for (int i = 0; i < 3; i++) { gd[i] += gdist/L*d[i] + gD[i]/L - dg*d[i]; g0[i] -= gd[i]; }
}

static void boundary(const oc_params* P, const int* I, double* v, unsigned* mask, double* vf) {
    unsigned m = 0; double gf = P->ground_friction;
    for (int d = 0; d < 3; d++) {
        if (I[d] < 3 && v[d] < 0) {
            m |= 1u << d;
            if (d != 1 || gf == 0) v[d] = 0;
            else if (gf < 10) {
                memcpy(vf, v, 24);
                double lin = v[1]+1e-30, vit[3] = {v[0]-I[0]*1e-30, v[1]-lin-I[1]*1e-30, v[2]-I[2]*1e-30};
                double lit = sqrt(vit[0]*vit[0]+vit[1]*vit[1]+vit[2]*vit[2]+1e-8), a = 1+gf*lin/lit, s = a > 0 ? a : 0;
                for (int i = 0; i < 3; i++) v[i] = s*(vit[i]+I[i]*1e-30);
                v[1] = 0;
            } else v[0] = v[1] = v[2] = 0;
        }
        if (I[d] > P->n_grid-3 && v[d] > 0) { m |= 1u << (4+d); v[d] = 0; }
    }
    *mask = m;
}
static void boundary_bwd(const oc_params* P, const int* I, unsigned mask, const double* vf, double* g) {
    double gf = P->ground_friction;
    for (int d = 2; d >= 0; d--) {
        if (mask & (1u << (4+d))) g[d] = 0;
        if (mask & (1u << d)) {
            if (d != 1 || gf == 0) g[d] = 0;
            else if (gf < 10) {
                double lin = vf[1]+1e-30, vit[3] = {vf[0]-I[0]*1e-30, vf[1]-lin-I[1]*1e-30, vf[2]-I[2]*1e-30};
                double lit = sqrt(vit[0]*vit[0]+vit[1]*vit[1]+vit[2]*vit[2]+1e-8), a = 1+gf*lin/lit, s = a > 0 ? a : 0;
                double gn[3] = {g[0], 0, g[2]}, gs = 0, gvit[3];
                for (int i = 0; i < 3; i++) { gs += gn[i]*(vit[i]+I[i]*1e-30); gvit[i] = s*gn[i]; }
                double ga = a > 0 ? gs : 0, glin = gf*ga/lit, glit = -gf*ga*lin/(lit*lit);
                for (int i = 0; i < 3; i++) gvit[i] += glit/lit*vit[i];
                glin -= gvit[1];
                g[0] = gvit[0]; g[1] = gvit[1]+glin; g[2] = gvit[2];
            } else g[0] = g[1] = g[2] = 0;
        }
    }
}

/* ------------------------------------------------------------------ one substep, forward.  grid: work array [G][8] doubles (in4 | out3 | pad) */
void oc_substep_fwd(const oc_params* P, const double* pose0, const double* pose1, const double* x, const double* v, const double* C, const double* F,
                    double* xo, double* vo, double* Co, double* Fo, double* grid) {
    const int n = P->n_grid; const long long G = (long long)n*n*n;
    #pragma omp parallel for
    for (long long i = 0; i < G*8; i++) grid[i] = 0;
    #pragma omp parallel for schedule(static, 256)
    for (int p = 0; p < P->n; p++) {
        double nF[9], aff[9], fx[3], w[3][3]; int b[3];
        p2g_particle(P, C+9*p, F+9*p, nF, aff, NULL);
        memcpy(Fo+9*p, nF, 72);
        stencil(P, x+3*p, b, fx, w);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double wt = w[i][0]*w[j][1]*w[k][2], dp[3] = {(i-fx[0])*P->dx, (j-fx[1])*P->dx, (k-fx[2])*P->dx};
            double* gnode = grid + 8*(((long long)(b[0]+i)*n + b[1]+j)*n + b[2]+k);
            for (int c = 0; c < 3; c++) {
                double val = wt*(P->p_mass*v[3*p+c] + aff[c*3]*dp[0] + aff[c*3+1]*dp[1] + aff[c*3+2]*dp[2]);
                #pragma omp atomic
                gnode[c] += val;
            }
            #pragma omp atomic
            gnode[3] += wt*P->p_mass;
        }
    }
    #pragma omp parallel for schedule(static, 1024)
    for (long long node = 0; node < G; node++) {
        double* gn = grid + 8*node;
        if (!(gn[3] > 1e-12)) continue;
        int I[3] = {(int)(node/((long long)n*n)), (int)((node/n)%n), (int)(node%n)};
        double gp[3] = {I[0]*P->dx, I[1]*P->dx, I[2]*P->dx}, vv[3], vf[3]; unsigned mask;
        for (int c = 0; c < 3; c++) vv[c] = (1/gn[3])*gn[c] + P->grav_dv[c];
        for (int k = 0; k < P->n_prim; k++) sphere_collide(P, k, pose0+8*k, pose1+8*k, gp, vv, 0, NULL, NULL, NULL);
        boundary(P, I, vv, &mask, vf);
        gn[4] = vv[0]; gn[5] = vv[1]; gn[6] = vv[2];
    }
    #pragma omp parallel for schedule(static, 256)
    for (int p = 0; p < P->n; p++) {
        double fx[3], w[3][3], nv[3] = {0, 0, 0}, nC[9] = {0}; int b[3];
        stencil(P, x+3*p, b, fx, w);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double wt = w[i][0]*w[j][1]*w[k][2], dp[3] = {i-fx[0], j-fx[1], k-fx[2]};
            const double* gv = grid + 8*(((long long)(b[0]+i)*n + b[1]+j)*n + b[2]+k) + 4;
            for (int c = 0; c < 3; c++) { nv[c] += wt*gv[c]; for (int d = 0; d < 3; d++) nC[c*3+d] += 4*P->inv_dx*wt*gv[c]*dp[d]; }
        }
        for (int c = 0; c < 3; c++) {
            double y = x[3*p+c] + P->dt*nv[c], hi = 1-3*P->dx; y = y < hi ? y : hi; xo[3*p+c] = 0 < y ? y : 0; vo[3*p+c] = nv[c];
        }
        memcpy(Co+9*p, nC, 72);
    }
}

/* ------------------------------------------------------------------ one substep, adjoint (substep_grad).  grid: [G][16] doubles work array */
void oc_substep_bwd(const oc_params* P, const double* pose0, const double* pose1, const double* x, const double* v, const double* C, const double* F,
                    const double* gxn, const double* gvn, const double* gCn, const double* gFn, double* gx, double* gv, double* gC, double* gF,
                    double* gpose0, double* gpose1, double* grid) {
    const int n = P->n_grid; const long long G = (long long)n*n*n;
    /* layout per node: [0..3] in4, [4..6] out, [7] pad, [8..10] g_out, [11] pad, [12..15] g_in */
    #pragma omp parallel for
    for (long long i = 0; i < G*16; i++) grid[i] = 0;
    memset(gpose0, 0, sizeof(double)*8*P->n_prim); memset(gpose1, 0, sizeof(double)*8*P->n_prim);
    /* recompute P2G + grid op */
    #pragma omp parallel for schedule(static, 256)
    for (int p = 0; p < P->n; p++) {
        double nF[9], aff[9], fx[3], w[3][3]; int b[3];
        p2g_particle(P, C+9*p, F+9*p, nF, aff, NULL);
        stencil(P, x+3*p, b, fx, w);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double wt = w[i][0]*w[j][1]*w[k][2], dp[3] = {(i-fx[0])*P->dx, (j-fx[1])*P->dx, (k-fx[2])*P->dx};
            double* gnode = grid + 16*(((long long)(b[0]+i)*n + b[1]+j)*n + b[2]+k);
            for (int c = 0; c < 3; c++) {
                double val = wt*(P->p_mass*v[3*p+c] + aff[c*3]*dp[0] + aff[c*3+1]*dp[1] + aff[c*3+2]*dp[2]);
                #pragma omp atomic
                gnode[c] += val;
            }
            #pragma omp atomic
            gnode[3] += wt*P->p_mass;
        }
    }
    #pragma omp parallel for schedule(static, 1024)
    for (long long node = 0; node < G; node++) {
        double* gn = grid + 16*node;
        if (!(gn[3] > 1e-12)) continue;
        int I[3] = {(int)(node/((long long)n*n)), (int)((node/n)%n), (int)(node%n)};
        double gp[3] = {I[0]*P->dx, I[1]*P->dx, I[2]*P->dx}, vv[3], vf[3]; unsigned mask;
        for (int c = 0; c < 3; c++) vv[c] = (1/gn[3])*gn[c] + P->grav_dv[c];
        for (int k = 0; k < P->n_prim; k++) sphere_collide(P, k, pose0+8*k, pose1+8*k, gp, vv, 0, NULL, NULL, NULL);
        boundary(P, I, vv, &mask, vf);
        gn[4] = vv[0]; gn[5] = vv[1]; gn[6] = vv[2];
    }
    /* g2p.grad */
    #pragma omp parallel for schedule(static, 256)
    for (int p = 0; p < P->n; p++) {
        double fx[3], w[3][3], nv[3] = {0, 0, 0}; int b[3];
        stencil(P, x+3*p, b, fx, w);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            const double* gvv = grid + 16*(((long long)(b[0]+i)*n + b[1]+j)*n + b[2]+k) + 4;
            double wt = w[i][0]*w[j][1]*w[k][2];
            for (int c = 0; c < 3; c++) nv[c] += wt*gvv[c];
        }
        double gy[3], gvt[3], gw[3][3] = {{0}}, gfx[3] = {0, 0, 0}, c4 = 4*P->inv_dx;
        for (int c = 0; c < 3; c++) {
            double y = x[3*p+c] + P->dt*nv[c], hi = 1-3*P->dx; int pm = y < hi; double z = pm ? y : hi;
            gy[c] = (pm && 0 < z) ? gxn[3*p+c] : 0; gvt[c] = gvn[3*p+c] + P->dt*gy[c];
        }
        const double* gCp = gCn + 9*p;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            double* gnode = grid + 16*(((long long)(b[0]+i)*n + b[1]+j)*n + b[2]+k);
            double wt = w[i][0]*w[j][1]*w[k][2], dp[3] = {i-fx[0], j-fx[1], k-fx[2]}, Cd[3], gwt = 0;
            for (int c = 0; c < 3; c++) Cd[c] = gCp[c*3]*dp[0] + gCp[c*3+1]*dp[1] + gCp[c*3+2]*dp[2];
            for (int c = 0; c < 3; c++) {
                double val = wt*(gvt[c] + c4*Cd[c]);
                #pragma omp atomic
                gnode[8+c] += val;
                gwt += gvt[c]*gnode[4+c] + c4*gnode[4+c]*Cd[c];
            }
            for (int d = 0; d < 3; d++) gfx[d] -= c4*wt*(gCp[d]*gnode[4] + gCp[3+d]*gnode[5] + gCp[6+d]*gnode[6]);
            gw[i][0] += gwt*w[j][1]*w[k][2]; gw[j][1] += gwt*w[i][0]*w[k][2]; gw[k][2] += gwt*w[i][0]*w[j][1];
        }
        for (int d = 0; d < 3; d++) {
            double g = gfx[d];
            for (int a = 0; a < 3; a++) g += gw[a][d]*dwt(a, fx[d]);
            gx[3*p+d] = gy[d] + g*P->inv_dx;
        }
    }
    /* grid_op.grad */
    int nt = 1;
    #ifdef _OPENMP
    nt = omp_get_max_threads();
    #endif
    double* gp_acc = (double*)calloc((size_t)nt*2*64, sizeof(double));
    #pragma omp parallel for schedule(static, 1024)
    for (long long node = 0; node < G; node++) {
        double* gn = grid + 16*node;
        if (!(gn[3] > 1e-12)) continue;
        int tid = 0;
        #ifdef _OPENMP
        tid = omp_get_thread_num();
        #endif
        int I[3] = {(int)(node/((long long)n*n)), (int)((node/n)%n), (int)(node%n)};
        double gp[3] = {I[0]*P->dx, I[1]*P->dx, I[2]*P->dx}, vv[3], vf[3], vstack[8][3]; unsigned mask;
        for (int c = 0; c < 3; c++) vv[c] = (1/gn[3])*gn[c] + P->grav_dv[c];
        for (int k = 0; k < P->n_prim; k++) { memcpy(vstack[k], vv, 24); sphere_collide(P, k, pose0+8*k, pose1+8*k, gp, vv, 0, NULL, NULL, NULL); }
        boundary(P, I, vv, &mask, vf);
        double g[3] = {gn[8], gn[9], gn[10]};
        boundary_bwd(P, I, mask, vf, g);
        for (int k = P->n_prim-1; k >= 0; k--)
            sphere_collide(P, k, pose0+8*k, pose1+8*k, gp, vstack[k], 1, g, gp_acc + ((size_t)tid*2)*64 + 8*k, gp_acc + ((size_t)tid*2+1)*64 + 8*k);
        double inv = 1/gn[3], ginv = g[0]*gn[0] + g[1]*gn[1] + g[2]*gn[2];
        gn[12] = inv*g[0]; gn[13] = inv*g[1]; gn[14] = inv*g[2]; gn[15] = -ginv*inv*inv;
    }
    for (int t = 0; t < nt; t++) for (int i = 0; i < 8*P->n_prim; i++) { gpose0[i] += gp_acc[((size_t)t*2)*64+i]; gpose1[i] += gp_acc[((size_t)t*2+1)*64+i]; }
    free(gp_acc);
    /* p2g.grad, svd_grad, compute_F_tmp.grad */
    #pragma omp parallel for schedule(static, 256)
    for (int p = 0; p < P->n; p++) {
        double nF[9], aff[9], fx[3], w[3][3], gvv[3] = {0, 0, 0}, gaff[9] = {0}, gw[3][3] = {{0}}, gfx[3] = {0, 0, 0}; int b[3];
        p2g_keep keep;
        p2g_particle(P, C+9*p, F+9*p, nF, aff, &keep);
        stencil(P, x+3*p, b, fx, w);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
            const double* a = grid + 16*(((long long)(b[0]+i)*n + b[1]+j)*n + b[2]+k) + 12;
            double wt = w[i][0]*w[j][1]*w[k][2], dp[3] = {(i-fx[0])*P->dx, (j-fx[1])*P->dx, (k-fx[2])*P->dx}, gwt = a[3]*P->p_mass;
            for (int c = 0; c < 3; c++) {
                gwt += a[c]*(P->p_mass*v[3*p+c] + aff[c*3]*dp[0] + aff[c*3+1]*dp[1] + aff[c*3+2]*dp[2]);
                gvv[c] += wt*a[c];
                for (int d = 0; d < 3; d++) gaff[c*3+d] += wt*a[c]*dp[d];
            }
            for (int d = 0; d < 3; d++) gfx[d] -= P->dx*wt*(aff[d]*a[0] + aff[3+d]*a[1] + aff[6+d]*a[2]);
            gw[i][0] += gwt*w[j][1]*w[k][2]; gw[j][1] += gwt*w[i][0]*w[k][2]; gw[k][2] += gwt*w[i][0]*w[j][1];
        }
        for (int d = 0; d < 3; d++) {
            double g = gfx[d];
            for (int a = 0; a < 3; a++) g += gw[a][d]*dwt(a, fx[d]);
            gx[3*p+d] += g*P->inv_dx; gv[3*p+d] = P->p_mass*gvv[d];
        }
        p2g_particle_bwd(P, C+9*p, F+9*p, &keep, gaff, gFn+9*p, gC+9*p, gF+9*p);
    }
}

void oc_set_threads(int n) {
    #ifdef _OPENMP
    omp_set_num_threads(n);
    #else
    (void)n;
    #endif
}

int oc_max_threads(void) {
    #ifdef _OPENMP
    return omp_get_max_threads();
    #else
    return 1;
    #endif
}
