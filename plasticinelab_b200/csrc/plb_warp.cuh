// Warp-level pre-reduction of the 27-node scatters (P2G and the G2P adjoint).
//
// After the spatial sort the 32 particles of a warp sit in a handful of cells, so their 27-node stencils coincide and
// per-particle atomics serialise at the L2 atomic unit (measured: 1M particles, p2g 230 us unsorted -> 439 us sorted).
// Every lane parks its contributions in a per-warp shared tile [nodes][33 columns]; lanes are grouped by base cell and
// for each group one lane per node sums the group's columns and issues ONE vector RED: atomics drop from
// 27 x 32 per warp to 27 x (#cells in the warp).
//
// Tile shape: [27][33] Vec4 (14.25 KB / warp), one flush per particle, lane q < 27 owns node q.  (A 9-node plane tile flushed after
// every stencil plane, 64-thread CTAs, a paired-cell flush and a one-kernel forward grid stage were measured in round 1 and
// lost at every size; they were removed in round 2, profiles/r1b_ab_results.md has the numbers.)
// Column 32 of every row (the padding that makes the transposed read conflict-free) is kept at zero and serves as the
// "no member" column: the member walk is unrolled by four with independent LDS.
//
// The code below uses only the small intrinsic layer (warp_ballot / warp_shfl / warp_shfl_down / warp_sync / ctz32), so
// that tests/host/warp_emul.hpp can run it on the CPU with 32 lock-stepped threads per warp (there is no GPU on the
// build box); on the device these are the CUDA warp intrinsics.
#pragma once
#include "plb_bodies.cuh"

namespace plb {

#if defined(__CUDACC__)
PLB_D unsigned warp_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
PLB_D int warp_shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
PLB_D float warp_shfl_down(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
PLB_D double warp_shfl_down(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
PLB_D float warp_shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
PLB_D double warp_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
PLB_D void warp_sync() { __syncwarp(); }
PLB_D int ctz32(unsigned g) { return __clz(__brev(g)); }          // 32 for g == 0
PLB_D void atomic_add_f64(double* dst, double v) { atomicAdd(dst, v); }
#endif
// (host: tests/host/warp_emul.hpp defines the same functions before this header is included)

// L2 prefetch of the line holding *p (no-op on the host).  The backward particle kernel of substep s uses it on the frame of
// substep s-1, which the NEXT kernel reads from HBM first thing (12 % of the fused backward kernel's stall samples sat on
// those first loads, profiles/r1b_move100k_ncu_stall_hotspots.txt).
#if defined(__CUDACC__)
PLB_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
inline void prefetch_l2(const void*) {}
#endif
// every plane of frame f at particle p except A0 (the fused backward kernel reads A0 of that frame itself)
template <class T> PLB_D void prefetch_frame_rest(const FramePtr<T>& f, int p) {
    prefetch_l2(f.A1 + p); prefetch_l2(f.A2 + p); prefetch_l2(f.a3 + p); prefetch_l2(f.a4 + p); prefetch_l2(f.a5 + p);
    prefetch_l2(f.B0 + p); prefetch_l2(f.B1 + p); prefetch_l2(f.b2 + p);
}

constexpr int kTileStride = 33;
constexpr int kNullCol = 32;
constexpr int kTileVec4 = 27 * kTileStride;

// base cell packed into one int, 10 bits per axis (n_grid <= 1024); negative = the lane carries no particle
PLB_HD int pack_cell(int i, int j, int k) { return (i << 20) | (j << 10) | k; }
template <class T> PLB_HD int cell_key(V3<T> x, T inv_dx) {
    return pack_cell((int)(x.x * inv_dx - T(0.5)), (int)(x.y * inv_dx - T(0.5)), (int)(x.z * inv_dx - T(0.5)));
}

// payload helpers: the tile carries Vec4 (momentum+mass, velocity adjoint) or a scalar (loss mass)
template <class T> PLB_HD void pay_zero(Vec4<T>& a) { a = mk4<T>(T(0), T(0), T(0), T(0)); }
template <class T> PLB_HD void pay_acc(Vec4<T>& a, const Vec4<T>& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
template <class T> PLB_D void pay_red(Vec4<T>* dst, const Vec4<T>& a) { scatter_add4(dst, a); }
PLB_HD void pay_zero(float& a) { a = 0.f; }
PLB_HD void pay_zero(double& a) { a = 0.0; }
PLB_HD void pay_acc(float& a, const float& v) { a += v; }
PLB_HD void pay_acc(double& a, const double& v) { a += v; }
PLB_D void pay_red(float* dst, const float& a) { scatter_add1(dst, a); }
PLB_D void pay_red(double* dst, const double& a) { scatter_add1(dst, a); }

// sum of the columns named by the bits of `g` of one tile row (row[kNullCol] must be zero)
template <class Pay> PLB_D Pay tile_row_sum(const Pay* row, unsigned g) {
    Pay a0, a1;
    pay_zero(a0); pay_zero(a1);
    while (g) {
        const int j0 = ctz32(g); g &= g - 1;
        const int j1 = ctz32(g); g &= g - 1;
        const int j2 = ctz32(g); g &= g - 1;
        const int j3 = ctz32(g); g &= g - 1;
        const Pay v0 = row[j0], v1 = row[j1], v2 = row[j2], v3 = row[j3];
        pay_acc(a0, v0); pay_acc(a1, v1); pay_acc(a0, v2); pay_acc(a1, v3);
    }
    pay_acc(a0, a1);
    return a0;
}

// ------------------------------------------------------------------------------------------------ direct halo (multi-GPU slabs)
// The neighbours' copies of the grid this kernel scatters into, mapped through CUDA IPC (NVLink peer memory).  A reduced
// contribution to a node whose plane lies in the zone shared with a neighbour is added to the neighbour's grid as well (one more
// RED, over NVLink), so that after the scatter kernels of both ranks the zone holds the full sums on both sides and no separate
// push / add pass is needed.  grid[side] == nullptr: no neighbour on that side (or single-GPU run: both null).
template <class Pay> struct PeerHalo {
    Pay* grid[2]; int lo[2], hi[2];
    PLB_HD bool any() const { return grid[0] != nullptr || grid[1] != nullptr; }
};
template <class Pay> PLB_HD PeerHalo<Pay> no_peers() { PeerHalo<Pay> p; p.grid[0] = p.grid[1] = nullptr; p.lo[0] = p.lo[1] = p.hi[0] = p.hi[1] = 0; return p; }
// plane = first grid index of the node; returns true if a remote RED was issued
template <class Pay> PLB_D bool pay_red_peers(const PeerHalo<Pay>& ph, long long node, int plane, const Pay& a) {
    bool sent = false;
#pragma unroll
    for (int side = 0; side < 2; side++)
        if (ph.grid[side] && plane >= ph.lo[side] && plane < ph.hi[side]) { pay_red(ph.grid[side] + node, a); sent = true; }
    return sent;
}

// ------------------------------------------------------------------------------------------------ full tile
template <class T> struct WarpTileScatter {
    Vec4<T>* tile;     // this warp's tile
    int lane;
    PLB_D void add(int slot, int, int, int, Vec4<T> v) const { tile[slot * kTileStride + lane] = v; }
    PLB_D void end_plane(int) const {}
};

// lane q < 27 zeroes the padding column of row q (the only lane that ever reads that row)
template <class Pay> PLB_D void tile_init(Pay* tile, int lane) {
    if (lane < 27) pay_zero(tile[lane * kTileStride + kNullCol]);
}

// key: packed base cell of this lane's particle (< 0: none).  All 32 lanes must call.  Loop over the distinct cells of
// the warp; lane q < 27 sums node q over the lanes of the cell and adds it to the grid with one (vector) RED.
// ph (optional): direct halo -- returns true if this lane issued a RED into a neighbour's grid
template <class Pay>
PLB_D bool warp_tile_flush(const Pay* tile, int lane, int key, int n_grid, Pay* grid, const PeerHalo<Pay>* ph = nullptr) {
    warp_sync();
    unsigned remaining = warp_ballot(key >= 0);
    const int oi = lane / 9, oj = (lane / 3) % 3, ok = lane % 3;       // node offset owned by this lane (lane < 27)
    const Pay* row = tile + (lane < 27 ? lane : 0) * kTileStride;
    bool sent = false;
    while (remaining) {
        const int leader = ctz32(remaining);
        const int lkey = warp_shfl(key, leader);
        const unsigned group = warp_ballot(key == lkey);
        remaining &= ~group;
        if (lane < 27) {
            const Pay acc = tile_row_sum(row, group);
            const long long node = node_index(n_grid, (lkey >> 20) + oi, ((lkey >> 10) & 1023) + oj, (lkey & 1023) + ok);
            pay_red(grid + node, acc);
            if (ph) sent = pay_red_peers(*ph, node, (lkey >> 20) + oi, acc) || sent;
        }
    }
    warp_sync();
    return sent;
}

// ------------------------------------------------------------------------------------------------ run-based flush
#if defined(__CUDA_ARCH__)
PLB_D Vec4<float> add4(Vec4<float> a, Vec4<float> b) {          // two packed FADD2 (sm_100)
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return mk4<float>(lo.x, lo.y, hi.x, hi.y);
}
#else
PLB_D Vec4<float> add4(Vec4<float> a, Vec4<float> b) { return mk4<float>(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
#endif
PLB_D Vec4<double> add4(Vec4<double> a, Vec4<double> b) { return mk4<double>(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// key: packed base cell of this lane's particle (< 0: none).  All 32 lanes must call.  A run = maximal stretch of consecutive
// lanes with one key (after the spatial sort a warp is a few runs; for an arbitrary order the result is still correct, the runs
// just get short); lane q < 27 walks the 32 columns of node q once, four LDS.128 at a time, and adds each run's sum to the grid
// with one vector RED.  Columns of lanes without a particle are read but land in a run of their own that is dropped.
// ph: direct halo -- returns true if this lane issued a RED into a neighbour's grid
template <class T>
PLB_D bool flush_runs(const Vec4<T>* tile, int lane, int key, int n_grid, Vec4<T>* grid, const PeerHalo<Vec4<T>>& ph) {
    bool sent = false;
    warp_sync();
    const int next = warp_shfl(key, (lane + 1) & 31);
    const unsigned ends = warp_ballot(lane == 31 || next != key);
    const int oi = lane / 9, oj = (lane / 3) % 3, ok = lane % 3;
    const Vec4<T>* row = tile + (lane < 27 ? lane : 0) * kTileStride;
    Vec4<T> acc = mk4<T>(T(0), T(0), T(0), T(0));
#pragma unroll 2
    for (int g = 0; g < 8; g++) {
        const Vec4<T> v0 = row[4 * g], v1 = row[4 * g + 1], v2 = row[4 * g + 2], v3 = row[4 * g + 3];
        const unsigned m = (ends >> (4 * g)) & 0xFu;          // warp-uniform
        if (m == 0u) {
            acc = add4(acc, add4(add4(v0, v1), add4(v2, v3)));
        } else {
            const Vec4<T> v[4] = {v0, v1, v2, v3};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                acc = add4(acc, v[j]);
                if ((m >> j) & 1u) {
                    const int rkey = warp_shfl(key, 4 * g + j);
                    if (lane < 27 && rkey >= 0) {
                        const long long node = node_index(n_grid, (rkey >> 20) + oi, ((rkey >> 10) & 1023) + oj, (rkey & 1023) + ok);
                        scatter_add4(grid + node, acc);
                        sent = pay_red_peers(ph, node, (rkey >> 20) + oi, acc) || sent;
                    }
                    acc = mk4<T>(T(0), T(0), T(0), T(0));
                }
            }
        }
    }
    warp_sync();
    return sent;
}

// mode 3 (default): runs of consecutive lanes (flush_runs: 297 instead of 615 warp instructions per flush in the fused backward
// kernel, profiles/r2c_summary.md); mode 0: per-cell groups (warp_tile_flush)
template <class Pay>
PLB_D bool warp_tile_flush_sel(Pay* tile, int lane, int key, int n_grid, Pay* grid, int mode, const PeerHalo<Pay>* ph = nullptr) {
    if (mode != 0) return ph ? flush_runs(tile, lane, key, n_grid, grid, *ph) : flush_runs(tile, lane, key, n_grid, grid, no_peers<Pay>());
    return warp_tile_flush(tile, lane, key, n_grid, grid, ph);
}


// ================================================================================================
// Thread-level scatter kernels: what ONE thread of a scatter kernel does, given its particle index, its lane and its warp's
// tile.  The __global__ wrappers in plb_kernels.cuh only derive (p, lane, tile) from the launch geometry; the host warp
// emulation (tests/host) calls the same functions.  Lanes past the end skip the math and only take part in the flush.
// ================================================================================================
namespace detail {
constexpr int kBlkShiftW = 2;          // 4 nodes per active-block edge (same constant as plb_kernels.cuh)
}

// flags[] of the (up to 8) 4^3 blocks touched by the stencil of a particle at x.  Plain byte stores, no test-before-set: a
// load of the flag would sit on the critical path (L2 latency, and the compiler cannot hoist it over the previous store); the
// stores of a warp to one flag coalesce.  Along an axis the stencil (base .. base+2) stays in one block unless base % 4 >= 2.
template <class T>
PLB_HD void mark_blocks(const SimConst<T>& P, V3<T> x, unsigned char* flags) {
    const int nbx = P.n_grid >> detail::kBlkShiftW;
    int lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const int b = (int)(x[d] * P.inv_dx - T(0.5));
        lo[d] = b >> detail::kBlkShiftW;
        hi[d] = (b + 2) >> detail::kBlkShiftW;
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if ((a && hi[0] == lo[0]) || (c && hi[1] == lo[1]) || (e && hi[2] == lo[2])) continue;      // same block as a = 0 / c = 0 / e = 0
                flags[((a ? hi[0] : lo[0]) * nbx + (c ? hi[1] : lo[1])) * nbx + (e ? hi[2] : lo[2])] = 1;
            }
}

// P2G of one substep
template <class T>
PLB_D void t_p2g(int p, int lane, Vec4<T>* tile, const SimConst<T>& P, const FramePtr<T>& fin, const FramePtr<T>& fout, bool store_F,
                 const Material<T>& mat, Vec4<T>* grid_in, unsigned char* flags, int flush_mode = 0, const SvdPtr<T>* svd_keep = nullptr) {
    const bool valid = p < P.n_particles;
    tile_init(tile, lane);
    int key = -1;
    if (valid) {
        WarpTileScatter<T> sc{tile, lane};
        p2g_body<T, WarpTileScatter<T>>(p, P, fin, fout, store_F, mat, sc, svd_keep);
        V3<T> x = load_x(fin, p);
        key = cell_key(x, P.inv_dx);
        if (flags) mark_blocks<T>(P, x, flags);
    }
    warp_tile_flush_sel(tile, lane, key, P.n_grid, grid_in, flush_mode);
}

// G2P of substep s (frame fin -> fmid) + P2G of substep s+1 (F from fmid, F' to fout), keyed on the ADVECTED position
template <class T>
PLB_D void t_g2p_p2g(int p, int lane, Vec4<T>* tile, const SimConst<T>& P, const FramePtr<T>& fin, const FramePtr<T>& fmid,
                     const FramePtr<T>& fout, const Material<T>& mat, const Vec4<T>* grid_out, Vec4<T>* grid_in, unsigned char* flags,
                     int flush_mode = 0, const SvdPtr<T>* svd_keep = nullptr) {
    const bool valid = p < P.n_particles;
    tile_init(tile, lane);
    int key = -1;
    if (valid) {
        M3<T> F = load_F(fmid, p);             // issued before the gather: the stores below may alias for the compiler
        T mu, lam, ys;
        load_material(P, mat, p, mu, lam, ys);
        V3<T> nx, nv; M3<T> nC;
        g2p_core<T>(P, load_x(fin, p), grid_out, nx, nv, nC);
        store_xvC(fmid, p, nx, nv, nC);
        M3<T> new_F;
        WarpTileScatter<T> sc{tile, lane};
        SvdRec<T> rec;
        p2g_core<T, WarpTileScatter<T>>(P, nx, nv, nC, F, mu, lam, ys, new_F, sc, svd_keep ? &rec : nullptr);
        store_F(fout, p, new_F);
        if (svd_keep) store_svd(*svd_keep, p, rec);
        key = cell_key(nx, P.inv_dx);
        if (flags) mark_blocks<T>(P, nx, flags);
    }
    warp_tile_flush_sel(tile, lane, key, P.n_grid, grid_in, flush_mode);
}

// g2p.grad of one substep (state frame fin; fnext = the frame G2P produced, or null pointers => recompute the gather sum)
template <class T>
PLB_D bool t_g2p_bwd(int p, int lane, Vec4<T>* tile, const SimConst<T>& P, const FramePtr<T>& fin, const FramePtr<T>* fnext,
                     const FramePtr<T>& adj_next, const FramePtr<T>& adj_cur, const Vec4<T>* grid_out, Vec4<T>* g_out, int flush_mode = 0,
                     const PeerHalo<Vec4<T>>* ph = nullptr) {
    const bool valid = p < P.n_particles;
    tile_init(tile, lane);
    int key = -1;
    if (valid) {
        V3<T> x = load_x(fin, p);
        WarpTileScatter<T> sc{tile, lane};
        V3<T> gxn, gvn; M3<T> gCn;
        load_xvC(adj_next, p, gxn, gvn, gCn);
        V3<T> gx;
        if (fnext) {
            Vec4<T> q0 = fnext->A0[p], q1 = fnext->A1[p];
            gx = g2p_bwd_core<T, WarpTileScatter<T>, true>(P, x, gxn, gvn, gCn, grid_out, sc, mk3<T>(q0.x, q0.y, q0.z), mk3<T>(q0.w, q1.x, q1.y));
        } else {
            gx = g2p_bwd_core<T, WarpTileScatter<T>, false>(P, x, gxn, gvn, gCn, grid_out, sc, zero3<T>(), zero3<T>());
        }
        adj_cur.A0[p] = mk4<T>(gx.x, gx.y, gx.z, T(0));
        key = cell_key(x, P.inv_dx);
    }
    return warp_tile_flush_sel(tile, lane, key, P.n_grid, g_out, flush_mode, ph);
}

// p2g.grad of substep s (frame fs) + g2p.grad of substep s-1 (frame fprev); the adjoint of (x,v,C)[s] stays in registers and
// the state (x,v)[s] this thread loaded anyway is what G2P(s-1) produced (clamp masks + gather sum come from it)
template <class T, bool kSvdGiven = false, bool kTwoPhase = false>
PLB_D bool t_p2g_bwd_g2p_bwd(int p, int lane, Vec4<T>* tile, const SimConst<T>& P, const FramePtr<T>& fs, const FramePtr<T>& fprev,
                             const FramePtr<T>& next, const FramePtr<T>& cur, const Material<T>& mat, const Vec4<T>* g_in,
                             const Vec4<T>* grid_out, Vec4<T>* g_out, int flush_mode = 0, const SvdPtr<T>* svd_kept = nullptr,
                             const PeerHalo<Vec4<T>>* ph = nullptr) {
    const bool valid = p < P.n_particles;
    if (valid) prefetch_frame_rest(fprev, p);          // for the next backward kernel (p2g.grad of substep s-1)
    tile_init(tile, lane);
    int key = -1;
    if (valid) {
        V3<T> x, v; M3<T> C;
        load_xvC(fs, p, x, v, C);
        M3<T> F = load_F(fs, p);
        T mu, lam, ys;
        load_material(P, mat, p, mu, lam, ys);
        Vec4<T> part = cur.A0[p];
        V3<T> gx, gv; M3<T> gC, gF;
        SvdRec<T> rec;
        if (kSvdGiven) rec = load_svd(*svd_kept, p);
        const M3<T> gF_next = kTwoPhase ? zeroM<T>() : load_F(next, p);
        pdl_wait();          // g_in (grid adjoint of this substep) is final; the particle loads above are already in flight
        p2g_bwd_core<T, kSvdGiven, kTwoPhase>(P, x, v, C, F, mu, lam, ys, g_in, gF_next, mk3<T>(part.x, part.y, part.z), gx, gv, gC, gF,
                                              &rec, &next, p, &fs);
        store_F(cur, p, gF);
        V3<T> xp = load_x(fprev, p);
        WarpTileScatter<T> sc{tile, lane};
        V3<T> gxp = g2p_bwd_core<T, WarpTileScatter<T>, true>(P, xp, gx, gv, gC, grid_out, sc, x, v);
        next.A0[p] = mk4<T>(gxp.x, gxp.y, gxp.z, T(0));
        key = cell_key(xp, P.inv_dx);
    } else {
        pdl_wait();
    }
    return warp_tile_flush_sel(tile, lane, key, P.n_grid, g_out, flush_mode, ph);
}

// mass-only scatter of the loss (scalar payload, full tile of scalars)
template <class T>
PLB_D void t_loss_mass(int p, int lane, T* tile, const SimConst<T>& P, const FramePtr<T>& fin, T* grid_mass) {
    tile_init(tile, lane);
    int key = -1;
    if (p < P.n_particles) {
        V3<T> x = load_x(fin, p);
        Stencil<T> st = make_stencil(x, P.inv_dx);
        key = pack_cell(st.b[0], st.b[1], st.b[2]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    tile[((i * 3 + j) * 3 + k) * kTileStride + lane] = st.w[i][0] * st.w[j][1] * st.w[k][2] * P.p_mass;
    }
    warp_tile_flush(tile, lane, key, P.n_grid, grid_mass);
}

// ================================================================================================
// grid_op.grad for one node per thread with the pose gradients of ONE primitive at a time in registers
// (k_grid_bwd_sparse_v2).  The array form (grid_bwd_body: PoseGrad g0[8], g1[8] per thread, reduced after the node) keeps
// 128 floats per thread in local memory; here each primitive's pair is warp-reduced right after its collide adjoint and
// added to the global pose gradient with double atomics from lane 0.  All 32 lanes must call (warp collectives inside the
// uniform loop over primitives); `act` = this thread has a node, `owned` = its pose gradients count on this rank (slab).
// ================================================================================================
template <class T>
PLB_D void reduce_pose_grad_T(const PoseGrad<T>& g, int lane, double* dst) {
    T vals[8] = {g.pos.x, g.pos.y, g.pos.z, g.rot.w, g.rot.x, g.rot.y, g.rot.z, g.gap};
#pragma unroll
    for (int c = 0; c < 8; c++) {
        T s = vals[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += warp_shfl_xor(s, o);
        if (lane == 0 && s != T(0)) atomic_add_f64(dst + c, (double)s);
    }
}

template <class T>
PLB_D void t_grid_bwd_node(bool act, bool owned, long long node, int lane, const SimConst<T>& P, const PrimSet<T>& prims, const Pose<T>* s0,
                           const Pose<T>* s1, Vec4<T>* grid_in, Vec4<T>* g_out, Vec4<T>* g_in, bool clear, double* prim_grad, int pf) {
    Vec4<T> in4 = mk4<T>(T(0), T(0), T(0), T(0)), go = in4;
    int ix = 0, iy = 0, iz = 0;
    if (act) {
        in4 = grid_in[node];
        go = g_out[node];
        const int n = P.n_grid;
        const unsigned un = (unsigned)node, nn = (unsigned)n, row = un / nn;        // n_grid <= 1024: node < 2^30, 32-bit divisions
        iz = (int)(un - row * nn); ix = (int)(row / nn); iy = (int)(row - (unsigned)ix * nn);
    }
    const bool live = act && (in4.w > T(1e-12));
    V3<T> vstack[PLB_MAX_PRIM];
    V3<T> g = zero3<T>(), gpos = zero3<T>(), vin = mk3<T>(in4.x, in4.y, in4.z);
    T inv_m = T(0);
    if (live) {
        inv_m = T(1) / in4.w;
        V3<T> v = mk3<T>(inv_m * in4.x + P.grav_dv[0], inv_m * in4.y + P.grav_dv[1], inv_m * in4.z + P.grav_dv[2]);
        gpos = mk3<T>(T(ix) * P.dx, T(iy) * P.dx, T(iz) * P.dx);
        for (int k = 0; k < P.n_prim; k++) {
            vstack[k] = v;
            bool taken;
            v = prim_collide(prims.s[k], s0[k], s1[k], gpos, v, P.dt, taken);
        }
        BoundaryTape<T> tape;
        boundary_forward<T>(P, ix, iy, iz, v, &tape);
        g = boundary_backward<T>(P, ix, iy, iz, tape, mk3<T>(go.x, go.y, go.z));
    }
    for (int k = P.n_prim - 1; k >= 0; k--) {                 // uniform trip count
        PoseGrad<T> g0, g1;
        g0.clear(); g1.clear();
        bool taken = false;
        if (live) g = prim_collide_bwd(prims.s[k], s0[k], s1[k], gpos, vstack[k], P.dt, g, g0, g1, taken);
        const bool contrib = taken && owned;
        if (warp_ballot(contrib) == 0u) continue;
        if (!contrib) { g0.clear(); g1.clear(); }
        reduce_pose_grad_T<T>(g0, lane, prim_grad + ((long long)pf * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
        reduce_pose_grad_T<T>(g1, lane, prim_grad + ((long long)(pf + 1) * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
    }
    if (act) {
        // v = inv_m * v_in + const ; inv_m = 1 / m
        g_in[node] = live ? mk4<T>(inv_m * g.x, inv_m * g.y, inv_m * g.z, -dot(g, vin) * inv_m * inv_m) : mk4<T>(T(0), T(0), T(0), T(0));
        if (clear) {
            if (in4.x != T(0) || in4.y != T(0) || in4.z != T(0) || in4.w != T(0)) grid_in[node] = mk4<T>(T(0), T(0), T(0), T(0));
            if (go.x != T(0) || go.y != T(0) || go.z != T(0)) g_out[node] = mk4<T>(T(0), T(0), T(0), T(0));
        }
    }
}

}  // namespace plb
