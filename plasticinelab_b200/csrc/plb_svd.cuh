// 3x3 SVD in registers and the reference's SVD adjoint.
//
// Forward: `ti.svd` (plb/engine/mpm_simulator.py:87-90) is Taichi's McAdams/Sifakis routine, which lives in
// the un-vendored taichi wheel.  What the hot path consumes is its CONVENTION -- det(U) = det(V) = +1, |sigma|
// descending, a negative determinant carried by the last singular value -- not its iteration scheme.  This
// implementation is a one-sided (Hestenes) Jacobi: rotate column pairs of A = F V until mutually orthogonal;
// it is accurate to working precision in both float and double and needs no square matrix products.
//
// Backward: `backward_svd` + `clamp` (mpm_simulator.py:97-115,143-151), literally: 1/clamp(s_j^2 - s_i^2, +-1e-6).
#pragma once
#include "plb_types.cuh"

namespace plb {

template <class T> struct SvdTol;
template <> struct SvdTol<float>  { static constexpr float  tol2 = 1e-14f; static constexpr int sweeps = 8; };
template <> struct SvdTol<double> { static constexpr double tol2 = 1e-31;  static constexpr int sweeps = 12; };

// warm (optional): V of the decomposition of a nearby matrix -- the same particle one substep earlier, F_tmp changes by
// O(dt |C|) per substep -- used as the starting Vacc.  A = F V0 then has nearly orthogonal columns and the iteration needs one
// rotating sweep (quadratic convergence) instead of three or four from the identity.  V0 is re-orthonormalised first
// (Gram-Schmidt, ~30 instructions), so rounding errors do not accumulate from substep to substep.  The result is an SVD of F
// to working precision either way; U S V^T, U V^T and the adjoint formula do not depend on which one (paired column signs).
template <class T>
PLB_HD void svd3(const M3<T>& F, M3<T>& U, V3<T>& sig, M3<T>& V, const M3<T>* warm = nullptr) {
    // a[c] = c-th column of A = F * Vacc, v[c] = c-th column of Vacc
    V3<T> a[3], v[3];
    if (warm) {
        V3<T> w0 = mk3<T>(warm->m[0][0], warm->m[1][0], warm->m[2][0]), w1 = mk3<T>(warm->m[0][1], warm->m[1][1], warm->m[2][1]);
        w0 = plb_rsqrt(dot(w0, w0)) * w0;
        w1 = w1 - dot(w0, w1) * w0;
        w1 = plb_rsqrt(dot(w1, w1)) * w1;
        v[0] = w0; v[1] = w1; v[2] = cross(w0, w1);               // det(V0) = +1
#pragma unroll
        for (int c = 0; c < 3; c++)
            a[c] = mk3<T>(F.m[0][0] * v[c].x + F.m[0][1] * v[c].y + F.m[0][2] * v[c].z, F.m[1][0] * v[c].x + F.m[1][1] * v[c].y + F.m[1][2] * v[c].z,
                          F.m[2][0] * v[c].x + F.m[2][1] * v[c].y + F.m[2][2] * v[c].z);
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            a[c] = mk3<T>(F.m[0][c], F.m[1][c], F.m[2][c]);
            v[c] = mk3<T>(c == 0 ? T(1) : T(0), c == 1 ? T(1) : T(0), c == 2 ? T(1) : T(0));
        }
    }
    // squared column norms, kept up to date by the rotations (alpha' = alpha - t gamma, beta' = beta + t gamma); they only
    // steer the convergence test -- the singular values are taken from freshly computed norms after the loop
    T nn[3] = {dot(a[0], a[0]), dot(a[1], a[1]), dot(a[2], a[2])};
    for (int sweep = 0; sweep < SvdTol<T>::sweeps; sweep++) {
        bool rotated = false;
#pragma unroll
        for (int pair = 0; pair < 3; pair++) {
            const int p = (pair == 2) ? 1 : 0;
            const int q = (pair == 0) ? 1 : 2;
            const T alpha = nn[p], beta = nn[q], gamma = dot(a[p], a[q]);
            if (gamma * gamma > SvdTol<T>::tol2 * alpha * beta) {
                rotated = true;
                // the angle may be approximate (see plb_rcp_fast); c is refined so that c^2 + s^2 = c^2 (1 + t^2) = 1 to rounding
                const T zeta = (beta - alpha) * plb_rcp_fast(T(2) * gamma);
                const T rt = T(1) + zeta * zeta;
                T t = plb_rcp_fast(plb_abs(zeta) + rt * plb_rsqrt_fast(rt));          // 1 / (|zeta| + sqrt(1 + zeta^2))
                t = (zeta >= T(0)) ? t : -t;
                const T c = plb_rsqrt(T(1) + t * t);
                const T s = c * t;
                V3<T> ap = c * a[p] - s * a[q], aq = s * a[p] + c * a[q];
                a[p] = ap; a[q] = aq;
                V3<T> vp = c * v[p] - s * v[q], vq = s * v[p] + c * v[q];
                v[p] = vp; v[q] = vq;
                nn[p] = alpha - t * gamma; nn[q] = beta + t * gamma;
            }
        }
        if (!rotated) break;
    }
    T n0 = dot(a[0], a[0]), n1 = dot(a[1], a[1]), n2 = dot(a[2], a[2]);
    // sort columns by descending norm (3-element network), keeping a and v paired
#define PLB_SWAPCOL(i, j, ni, nj) { V3<T> ta = a[i]; a[i] = a[j]; a[j] = ta; V3<T> tv = v[i]; v[i] = v[j]; v[j] = tv; T tn = ni; ni = nj; nj = tn; }
    if (n0 < n1) PLB_SWAPCOL(0, 1, n0, n1)
    if (n1 < n2) PLB_SWAPCOL(1, 2, n1, n2)
    if (n0 < n1) PLB_SWAPCOL(0, 1, n0, n1)
#undef PLB_SWAPCOL
    // det(V) = +1
    if (dot(v[0], cross(v[1], v[2])) < T(0)) { v[2] = -v[2]; a[2] = -a[2]; }
    const T tiny2 = sizeof(T) == 4 ? T(1e-36) : T(1e-60);          // squared norm below which a column counts as zero
    const bool ok0 = n0 > tiny2, ok1 = n1 > tiny2;
    const T r0 = ok0 ? plb_rsqrt(n0) : T(0), r1 = ok1 ? plb_rsqrt(n1) : T(0);
    T s0 = n0 * r0, s1 = n1 * r1;               // sqrt(n) = n / sqrt(n)
    V3<T> u0 = ok0 ? r0 * a[0] : mk3<T>(T(1), T(0), T(0));
    V3<T> u1;
    if (ok1) {
        u1 = r1 * a[1];
    } else {  // rank <= 1: any unit vector orthogonal to u0
        V3<T> e = (plb_abs(u0.x) < T(0.9)) ? mk3<T>(T(1), T(0), T(0)) : mk3<T>(T(0), T(1), T(0));
        V3<T> w = cross(u0, e);
        u1 = (T(1) / plb_sqrt(dot(w, w))) * w;
    }
    V3<T> u2 = cross(u0, u1);             // det(U) = +1 by construction
    T s2 = dot(u2, a[2]);                 // signed
    sig = mk3<T>(s0, s1, s2);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        U.m[r][0] = u0[r]; U.m[r][1] = u1[r]; U.m[r][2] = u2[r];
        V.m[r][0] = v[0][r]; V.m[r][1] = v[1][r]; V.m[r][2] = v[2][r];
    }
}

// mpm_simulator.py:143-151
template <class T> PLB_HD T svd_clamp(T a) {
    if (a >= T(0)) return a > T(1e-6) ? a : T(1e-6);
    return a < T(-1e-6) ? a : T(-1e-6);
}

// Returns dL/dF given dL/dU, dL/d(sig diagonal), dL/dV (mpm_simulator.py:97-115; only the diagonal of the
// sigma adjoint can be non-zero on this path because p2g reads sig[i,i] only).
template <class T>
PLB_HD M3<T> svd3_backward(const M3<T>& gU, V3<T> gsig, const M3<T>& gV, const M3<T>& U, V3<T> sig, const M3<T>& V) {
    M3<T> UtgU = mTm(U, gU);        // U^T gU   ;  gU^T U is its transpose
    M3<T> VtgV = mTm(V, gV);
    T s2[3] = {sig.x * sig.x, sig.y * sig.y, sig.z * sig.z};
    T sg[3] = {sig.x, sig.y, sig.z};
    // Fm[i][j] = 1 / clamp(s_j^2 - s_i^2): three reciprocals serve the six off-diagonal entries, because clamp is odd except
    // at 0 (clamp(+-0) = +1e-6 both ways): Fm[j][i] = -Fm[i][j] unless s_j^2 == s_i^2 exactly, where both are +1e6
    M3<T> Fm = zeroM<T>();
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i + 1; j < 3; j++) {
            const T d = s2[j] - s2[i];
            const T f = plb_rcp_nr(svd_clamp(d));
            Fm.m[i][j] = f;
            Fm.m[j][i] = (d == T(0)) ? f : -f;
        }
    M3<T> inner;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (i == j) {
                inner.m[i][j] = gsig[i];
            } else {
                T au = UtgU.m[i][j] - UtgU.m[j][i];
                T av = VtgV.m[i][j] - VtgV.m[j][i];
                inner.m[i][j] = Fm.m[i][j] * au * sg[j] + sg[i] * Fm.m[i][j] * av;
            }
        }
    return mmT(mm(U, inner), V);
}

}  // namespace plb
