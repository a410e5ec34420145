// Small register-resident linear algebra for the MPM kernels (sm_100a).
//
// Everything here is `PLB_HD` (host+device) and templated on the scalar type so that
//   * the float instantiation is the production path,
//   * the double instantiation is the parity mode (the reference is float64-only,
//     plb/engine/mpm_simulator.py:8), and
//   * tests/host/emul.cpp can drive the very same per-particle / per-node math on the CPU
//     (there is no GPU on the build box) against the float64 oracle.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PLB_HD __host__ __device__ __forceinline__
#define PLB_D __device__ __forceinline__
// out-of-line on the device: the shape switches of the non-spherical primitives, so that the grid kernels (a handful of
// resident warps, instruction-fetch bound) keep a compact hot path instead of ~25 inlined copies of every shape
#define PLB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define PLB_HD inline
#define PLB_D inline
#define PLB_HD_NOINLINE inline
#endif

namespace plb {

// Programmatic dependent launch (sm_90+).  Inside the env-step graphs consecutive kernels are chained with programmatic edges
// (plb_engine.cu, launch_k): a kernel calls pdl_launch() once its dependent may become resident (the dependent then sets up --
// poses, block list, particle loads -- while this kernel's last wave drains) and pdl_wait() before it first touches anything
// the kernels ahead of it in the chain produce.  A grid kernel calls pdl_launch() only AFTER its own pdl_wait(), so the
// particle kernel behind it never starts before the particle kernel ahead of it has completed (its frames are final).
// Both are no-ops in a kernel launched without the attribute (and on the host).
#if defined(__CUDA_ARCH__)
PLB_D void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PLB_D void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
PLB_D void pdl_wait() {}
PLB_D void pdl_launch() {}
#endif

template <class T> struct V3 {
    T x, y, z;
    PLB_HD T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    PLB_HD const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> struct Q4 { T w, x, y, z; };          // quaternion (w, x, y, z)
template <class T> struct M3 { T m[3][3]; };              // row major, m[row][col]

template <class T> PLB_HD V3<T> mk3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <class T> PLB_HD V3<T> zero3() { return mk3<T>(T(0), T(0), T(0)); }
template <class T> PLB_HD V3<T> operator+(V3<T> a, V3<T> b) { return mk3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> PLB_HD V3<T> operator-(V3<T> a, V3<T> b) { return mk3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> PLB_HD V3<T> operator-(V3<T> a) { return mk3<T>(-a.x, -a.y, -a.z); }
template <class T> PLB_HD V3<T> operator*(T s, V3<T> a) { return mk3<T>(s * a.x, s * a.y, s * a.z); }
template <class T> PLB_HD V3<T> operator*(V3<T> a, T s) { return mk3<T>(s * a.x, s * a.y, s * a.z); }
template <class T> PLB_HD V3<T>& operator+=(V3<T>& a, V3<T> b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
template <class T> PLB_HD V3<T>& operator-=(V3<T>& a, V3<T> b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
template <class T> PLB_HD V3<T> fma3(T s, V3<T> a, V3<T> c) { return mk3<T>(s * a.x + c.x, s * a.y + c.y, s * a.z + c.z); }   // s a + c
template <class T> PLB_HD T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> PLB_HD V3<T> cross(V3<T> a, V3<T> b) {
    return mk3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

template <class T> PLB_HD M3<T> zeroM() {
    M3<T> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = T(0);
    return r;
}
template <class T> PLB_HD M3<T> identM() { M3<T> r = zeroM<T>(); r.m[0][0] = r.m[1][1] = r.m[2][2] = T(1); return r; }
template <class T> PLB_HD M3<T> operator+(const M3<T>& a, const M3<T>& b) {
    M3<T> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j];
    return r;
}
template <class T> PLB_HD M3<T> operator-(const M3<T>& a, const M3<T>& b) {
    M3<T> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j];
    return r;
}
template <class T> PLB_HD M3<T> operator*(T s, const M3<T>& a) {
    M3<T> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j];
    return r;
}
template <class T> PLB_HD M3<T>& operator+=(M3<T>& a, const M3<T>& b) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) a.m[i][j] += b.m[i][j];
    return a;
}
// a * b
template <class T> PLB_HD M3<T> mm(const M3<T>& a, const M3<T>& b) {
    M3<T> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
// a * b^T
template <class T> PLB_HD M3<T> mmT(const M3<T>& a, const M3<T>& b) {
    M3<T> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[j][0] + a.m[i][1] * b.m[j][1] + a.m[i][2] * b.m[j][2];
    return r;
}
// a^T * b
template <class T> PLB_HD M3<T> mTm(const M3<T>& a, const M3<T>& b) {
    M3<T> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[0][i] * b.m[0][j] + a.m[1][i] * b.m[1][j] + a.m[2][i] * b.m[2][j];
    return r;
}
template <class T> PLB_HD M3<T> transpose(const M3<T>& a) {
    M3<T> r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
    return r;
}
template <class T> PLB_HD V3<T> mv(const M3<T>& a, V3<T> v) {
    return mk3<T>(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
                  a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
                  a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
template <class T> PLB_HD V3<T> mTv(const M3<T>& a, V3<T> v) {
    return mk3<T>(a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z,
                  a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z,
                  a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z);
}
template <class T> PLB_HD T det(const M3<T>& a) {
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1])
         - a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0])
         + a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}
// d det(a) / d a  (cofactor matrix)
template <class T> PLB_HD M3<T> cofactor(const M3<T>& a) {
    M3<T> c;
    c.m[0][0] = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
    c.m[0][1] = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
    c.m[0][2] = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
    c.m[1][0] = a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2];
    c.m[1][1] = a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0];
    c.m[1][2] = a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1];
    c.m[2][0] = a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1];
    c.m[2][1] = a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2];
    c.m[2][2] = a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0];
    return c;
}
// outer product a b^T
template <class T> PLB_HD M3<T> outer(V3<T> a, V3<T> b) {
    M3<T> r;
    r.m[0][0] = a.x * b.x; r.m[0][1] = a.x * b.y; r.m[0][2] = a.x * b.z;
    r.m[1][0] = a.y * b.x; r.m[1][1] = a.y * b.y; r.m[1][2] = a.y * b.z;
    r.m[2][0] = a.z * b.x; r.m[2][1] = a.z * b.y; r.m[2][2] = a.z * b.z;
    return r;
}

// ---- scalar helpers with one spelling for float and double
PLB_HD float  plb_sqrt(float x) { return sqrtf(x); }
PLB_HD double plb_sqrt(double x) { return sqrt(x); }
PLB_HD float  plb_exp(float x) { return expf(x); }
PLB_HD double plb_exp(double x) { return exp(x); }
PLB_HD float  plb_log(float x) { return logf(x); }
PLB_HD double plb_log(double x) { return log(x); }
PLB_HD float  plb_abs(float x) { return fabsf(x); }
PLB_HD double plb_abs(double x) { return fabs(x); }
PLB_HD float  plb_sin(float x) { return sinf(x); }
PLB_HD double plb_sin(double x) { return sin(x); }
PLB_HD float  plb_cos(float x) { return cosf(x); }
PLB_HD double plb_cos(double x) { return cos(x); }

// Reciprocal and reciprocal square root used inside the Jacobi SVD.  double / host: exact IEEE forms.  float on the
// device: MUFU approximation + one Newton step (<= 1 ulp), a third of the instructions of the IEEE div/sqrt sequences.
PLB_HD double plb_rcp(double x) { return 1.0 / x; }
PLB_HD double plb_rsqrt(double x) { return 1.0 / sqrt(x); }
#if defined(__CUDA_ARCH__)
PLB_HD float plb_rcp(float x) { float r = __frcp_rn(x); return r; }
PLB_HD float plb_rsqrt(float x) { float r = rsqrtf(x); return r * (1.5f - 0.5f * x * r * r); }
#else
PLB_HD float plb_rcp(float x) { return 1.0f / x; }
PLB_HD float plb_rsqrt(float x) { return 1.0f / sqrtf(x); }
#endif

// Raw MUFU approximations (float on the device, ~1e-7 relative, no special-case handling) for quantities whose error only
// perturbs an iteration that corrects itself (the Jacobi rotation ANGLE: any angle close to the optimal one still converges,
// the rotation itself is built from a refined c with s = c t, so it stays orthogonal to rounding), and a 2-instruction refined
// reciprocal for the places that used an IEEE division of well-scaled positive values.  double / host: exact.
PLB_HD double plb_rcp_fast(double x) { return 1.0 / x; }
PLB_HD double plb_rsqrt_fast(double x) { return 1.0 / sqrt(x); }
PLB_HD double plb_rcp_nr(double x) { return 1.0 / x; }
#if defined(__CUDA_ARCH__)
PLB_HD float plb_rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
PLB_HD float plb_rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
PLB_HD float plb_rcp_nr(float x) { float r = plb_rcp_fast(x); return r * (2.0f - x * r); }       // <= 1 ulp for normal x
#else
PLB_HD float plb_rcp_fast(float x) { return 1.0f / x; }
PLB_HD float plb_rsqrt_fast(float x) { return 1.0f / sqrtf(x); }
PLB_HD float plb_rcp_nr(float x) { return 1.0f / x; }
#endif

// ti.max / ti.min value semantics and the gradient routing of Taichi's autodiff:
//   max(a,b): d/da = [b < a], d/db = 1 - [b < a];   min(a,b): d/da = [a < b], d/db = 1 - [a < b].
template <class T> PLB_HD T tmax(T a, T b) { return (b < a) ? a : b; }
template <class T> PLB_HD T tmin(T a, T b) { return (a < b) ? a : b; }

// Optimisation barrier for one value: the compiler may not assume it equals anything it computed before (used to make a
// deliberately repeated computation really repeat instead of keeping its results alive in registers).
#if defined(__CUDA_ARCH__)
PLB_HD void opaque(float& v) { asm volatile("" : "+f"(v)); }
PLB_HD void opaque(double& v) { asm volatile("" : "+d"(v)); }
#else
PLB_HD void opaque(float&) {}
PLB_HD void opaque(double&) {}
#endif

// ---- 4-wide storage vector: 16 B (float) / 32 B (double), the unit of the particle planes and grids
template <class T> struct alignas(4 * sizeof(T)) Vec4 { T x, y, z, w; };
template <class T> PLB_HD Vec4<T> mk4(T x, T y, T z, T w) { Vec4<T> r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

}  // namespace plb
