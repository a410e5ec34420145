// Chunked particle kernels with TMA-loaded grid tiles (sm_100a).
//
// Particles are sorted by (4^3 grid block, cell) when a state is installed (plb_sort_particles), so the particles of one grid
// block are one contiguous range.  That range is cut into CHUNKS of at most 128 particles; one CTA owns one chunk.  All stencils
// of a chunk lie inside the 8^3-node window that starts one node below the block (block = 4 cells, stencil = 3 nodes, one cell
// of drift allowed either way since the sort), so the CTA fetches that window of the grid ONCE with a single TMA box copy
// (cp.async.bulk.tensor.4d: 4 scalars x 8 x 8 x 8 nodes = 8 KB in float32, completion on an mbarrier, out-of-range nodes
// zero-filled by the hardware) and every 27-node gather -- G2P, p2g.grad, g2p.grad -- becomes 27 shared-memory loads with
// compile-time offsets from one base address instead of 27 dependent L2 round trips with 64-bit index arithmetic
// (reference: g2p `plb/engine/mpm_simulator.py:223-242`, p2g.grad / g2p.grad `:260-278`).  A particle that drifted out of the
// window (faster than one cell per env step) reads the dense grid through the same `GridView` -- a generic pointer with the
// global strides -- so correctness never depends on the sort being fresh.
//
// The 27-node scatters (P2G `:157-184`, g2p.grad) keep the per-warp transposed tile of plb_warp.cuh, with a run-based flush:
// inside a chunk the lanes of a warp are sorted by cell, so equal cells are runs of consecutive lanes; lane q < 27 walks the
// 32 columns of node q once, four LDS.128 at a time, and issues one vector RED per run.  A tile is busy only while its warp
// parks and flushes (~10 % of the kernel), so the four warps of a CTA SHARE TWO tiles: warps 0/1 scatter first and then
// arrive at a named barrier that warps 2/3 wait on before they park.  The TMA windows alias the tiles (the gathers are
// finished, CTA-wide, before the first contribution is parked): 28.5 KB of shared memory per CTA instead of 57 KB, so the
// resident warps per SM are limited by registers (5 / 4 CTAs forward / backward), not by shared memory (3 CTAs).
#pragma once
#include <cuda.h>          // CUtensorMap (type only: the encoder is fetched at run time with cudaGetDriverEntryPoint, no libcuda link)
#include "plb_warp.cuh"

namespace plb {

// one CTA's work: `count` (<= kBlock) consecutive particles starting at `start`, all sorted into grid block (bi, bj, bk);
// origin = first node of the block's 8^3 gather window, packed (o0 + 1) << 20 | (o1 + 1) << 10 | (o2 + 1)  (o = 4 b - 1 >= -1)
struct Chunk { int origin, start, count, pad; };
struct ChunkTable { const Chunk* chunks; const int* n_chunks; };  // device-resident, rebuilt in place by every sort

constexpr int kTileEdge = 8;
constexpr int kTileNodes = kTileEdge * kTileEdge * kTileEdge;
constexpr int kChunk = 128;
// scatter tiles of a CTA: 2 = shared by warp pairs (forward kernels: shared memory would otherwise cap the resident CTAs below the
// register limit), 4 = one per warp (backward kernels: registers cap them at 4 CTAs per SM anyway, so sharing would only add waits)

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------ mbarrier + TMA (PTX)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    // (with a suspend-time hint: the waiting warp sleeps in the barrier unit instead of re-issuing the poll -- the polls of the
    //  tile hand-off were 8.7 % of the forward kernel's issued instructions, profiles/r2c_summary.md)
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase), "r"(0x989680u) : "memory");
}
// box copy of a 4-D tensor (component, k, j, i) into shared memory; completion (bytes) is signalled on `bar`
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, unsigned long long* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Scatter-tile protocol of warp w (0..3) with kTiles tiles per CTA: tile w % kTiles; with 2 tiles warps 2/3 wait until warps
// 0/1 have flushed.  The hand-off is an mbarrier per tile (hand[0..1], arrival count 1, initialised with the TMA barrier): the
// first version used named barriers (bar.sync / bar.arrive 1, 2), and the 16 hardware barriers of an SM then capped the kernel
// at 4 resident CTAs (ncu "Block Limit Barriers 4") -- below the 5 the registers allow, which is the point of sharing tiles.
// (the hand-off lasts ~600 warp instructions of the other warp pair: polling it back to back took 14.6 % of the forward kernel's issue
//  slots, profiles/r2j_slab1m_fwd_chunk_by_source.txt -- the waiting warp sleeps between polls instead)
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned phase) {
    unsigned ok = 0;
    for (;;) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        if (ok) break;
        __nanosleep(160);
    }
}
template <class T, int kTiles> __device__ __forceinline__ Vec4<T>* shared_tile_acquire(unsigned char* smem_raw, unsigned long long* hand, int w) {
    if (kTiles == 2 && w >= 2) mbar_wait_backoff(&hand[w & 1], 0);
    return reinterpret_cast<Vec4<T>*>(smem_raw) + (w % kTiles) * kTileVec4;
}
template <int kTiles> __device__ __forceinline__ void shared_tile_release(unsigned long long* hand, int w, int lane) {
    if (kTiles == 2 && w < 2) {
        __syncwarp();                                   // every lane's tile reads have returned (their values were consumed)
        if (lane == 0) mbar_arrive(&hand[w & 1]);
    }
}

__device__ __forceinline__ void tile_origin(const Chunk& ch, int o[3]) {
    o[0] = (ch.origin >> 20) - 1; o[1] = ((ch.origin >> 10) & 1023) - 1; o[2] = (ch.origin & 1023) - 1;
}
// the stencil window of base node b: inside the CTA's tile if it fits (b - o in [0, 5] per axis), else the dense grid
template <class T>
__device__ __forceinline__ GridView<T> pick_view(const Vec4<T>* tile, const int o[3], const Vec4<T>* grid, int n, const int b[3]) {
    const int l0 = b[0] - o[0], l1 = b[1] - o[1], l2 = b[2] - o[2];
    GridView<T> v;
    if ((unsigned)l0 <= 5u && (unsigned)l1 <= 5u && (unsigned)l2 <= 5u) {
        v.p = tile + ((l0 * kTileEdge + l1) * kTileEdge + l2); v.si = kTileEdge * kTileEdge; v.sj = kTileEdge;
    } else {
        v.p = grid + node_index(n, b[0], b[1], b[2]); v.si = n * n; v.sj = n;
    }
    return v;
}

// (the run-based flush of the scatter tiles, flush_runs, lives in plb_warp.cuh: the per-warp kernels use it too)

// ------------------------------------------------------------------------------------------------ forward
// kMode = FWD_P2G: P2G of the first substep of a graph (frame s_in -> F' into s_out, scatter)
//         FWD_G2P | FWD_P2G: G2P of substep s (x of s_in -> x', v', C' into s_mid) + P2G of substep s+1 (F of s_mid -> F' into s_out)
//         FWD_G2P: G2P of the last substep (x of s_in -> s_out)
enum { FWD_G2P = 1, FWD_P2G = 2 };
constexpr int kFwdTiles = 2, kBwdTiles = 4;
template <class T, int kMode, int kMinB>
__global__ void __launch_bounds__(kBlock, kMinB)
k_fwd_chunk(const __grid_constant__ CUtensorMap tm_out, SimConst<T> P, T* frames, long long n_pad, SlotRef s_in, SlotRef s_mid, SlotRef s_out,
            Material<T> mat, ChunkTable sg, const Vec4<T>* grid_out, Vec4<T>* grid_in, unsigned char* flags, T* svd_base, int svd_warm,
            PeerHalo<Vec4<T>> ph, HaloOut ho) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[3];          // [0] TMA window, [1..2] scatter-tile hand-off
    unsigned long long& mbar = bars[0];
    pdl_launch();                                       // (the grid kernel behind this one waits for its completion itself)
    const int n_chunks = *sg.n_chunks;
    const Chunk ch = sg.chunks[blockIdx.x];             // (the table has room for the whole launch grid: both loads are in flight together)
    if ((int)blockIdx.x >= n_chunks) { halo_publish_scatter(ho, false); return; }
    const int tid = threadIdx.x, lane = tid & 31;
    // warm start of the Jacobi SVD from V of the previous substep (record of slot s_in, written by the previous kernel)
    const bool warm = kMode == (FWD_G2P | FWD_P2G) && svd_base != nullptr && svd_warm != 0;
    Vec4<T>* gtile = reinterpret_cast<Vec4<T>*>(smem_raw);                                   // TMA window (aliases the scatter tiles)
    int o[3];
    tile_origin(ch, o);
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        if (kMode & FWD_P2G) { mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); }
    }
    __syncthreads();
    const bool valid = tid < ch.count;
    const int p = ch.start + tid;
    const FramePtr<T> fin = frame_at(frames, s_in.get(), n_pad);
    const FramePtr<T> fmid = frame_at(frames, s_mid.get(), n_pad);
    const FramePtr<T> fout = frame_at(frames, s_out.get(), n_pad);
    V3<T> x = zero3<T>(), v = zero3<T>();
    M3<T> C = zeroM<T>(), F = zeroM<T>();
    T mu = T(0), lam = T(0), ys = T(0);
    M3<T> V0 = zeroM<T>();
    if (valid) {
        // (everything that does not depend on the grid is in flight while the TMA copy lands)
        if (warm) V0 = load_svd_V(svd_at(svd_base, s_in.get(), n_pad), p);
        if (kMode == FWD_P2G) {
            load_xvC(fin, p, x, v, C);
            F = load_F(fin, p);
        } else {
            x = load_x(fin, p);
            if (kMode & FWD_P2G) F = load_F(fmid, p);
        }
        if (kMode & FWD_P2G) load_material(P, mat, p, mu, lam, ys);
    }
    // everything above reads what the particle kernel two launches back produced (complete: the grid kernel between the two
    // released this launch only after its own wait) and is in flight now; grid_out / grid_in belong to the grid kernel ahead
    pdl_wait();
    if (kMode & FWD_G2P) {
        if (tid == 0) {
            mbar_expect_tx(&mbar, (unsigned)(kTileNodes * sizeof(Vec4<T>)));
            tma_load_4d(gtile, &tm_out, &mbar, 0, o[2], o[1], o[0]);
        }
        mbar_wait(&mbar, 0);
        if (valid) {
            const Stencil<T> st = make_stencil(x, P.inv_dx);
            const GridView<T> gv = pick_view(gtile, o, grid_out, P.n_grid, st.b);
            V3<T> nx, nv; M3<T> nC;
            g2p_core<T>(P, x, st, gv, nx, nv, nC);
            store_xvC((kMode & FWD_P2G) ? fmid : fout, p, nx, nv, nC);
            x = nx; v = nv; C = nC;
        }
    }
    if (kMode & FWD_P2G) {
        M3<T> affine = zeroM<T>();
        Stencil<T> st;
        int key = -1;
        if (valid) {
            M3<T> new_F;
            SvdRec<T> rec;
            p2g_particle<T>(P, C, F, mu, lam, ys, new_F, affine, nullptr, svd_base ? &rec : nullptr, warm ? &V0 : nullptr);
            store_F(fout, p, new_F);
            if (svd_base) store_svd(svd_at(svd_base, (kMode & FWD_G2P) ? s_mid.get() : s_in.get(), n_pad), p, rec);
            st = make_stencil(x, P.inv_dx);
            key = pack_cell(st.b[0], st.b[1], st.b[2]);
            if (flags) mark_blocks<T>(P, x, flags);
        }
        if (kMode & FWD_G2P) __syncthreads();          // every warp is done with the window: the scatter tiles may overwrite it
        Vec4<T>* stile = shared_tile_acquire<T, kFwdTiles>(smem_raw, &bars[1], tid >> 5);
        if (valid) {
            WarpTileScatter<T> sc{stile, lane};
            p2g_scatter<T>(P, st, v, affine, sc);
        }
        const bool sent = flush_runs<T>(stile, lane, key, P.n_grid, grid_in, ph);
        shared_tile_release<kFwdTiles>(&bars[1], tid >> 5, lane);
        halo_publish_scatter(ho, sent);
    }
}

// ------------------------------------------------------------------------------------------------ backward
// kMode = BWD_G2P: g2p.grad of the last substep of a graph (state frame s_prev; adjoint of (x, v, C) of its successor in adj_next;
//                  clamp masks / gather sum from frame s_s = s_prev + 1 when next_ok) -> partial x-adjoint into adj_cur.A0
//         BWD_P2G | BWD_G2P: p2g.grad of substep s (frame s_s) + g2p.grad of substep s-1 (frame s_prev): the adjoint of (x, v, C)[s] stays
//                  in registers, adjoint of F[s] -> adj_cur, partial x-adjoint of frame s-1 -> adj_next.A0 (the buffers then swap roles)
//         BWD_P2G: p2g.grad of the first substep (frame s_s) -> full adjoint frame in adj_cur
enum { BWD_P2G = 1, BWD_G2P = 2 };
template <class T, int kMode, int kMinB, bool kSvd>
__global__ void __launch_bounds__(kBlock, kMinB)
k_bwd_chunk(const __grid_constant__ CUtensorMap tm_gin, const __grid_constant__ CUtensorMap tm_out, SimConst<T> P, T* frames, long long n_pad,
            SlotRef s_s, SlotRef s_prev, int next_ok, T* adj_next, T* adj_cur, Material<T> mat, ChunkTable sg,
            const Vec4<T>* g_in, const Vec4<T>* grid_out, Vec4<T>* g_out, T* svd_base) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar;
    constexpr bool kTwoPhase = kSvd && kMinB >= 4;
    const int n_chunks = *sg.n_chunks;
    const Chunk ch = sg.chunks[blockIdx.x];
    if ((int)blockIdx.x >= n_chunks) return;
    const int tid = threadIdx.x, lane = tid & 31;
    Vec4<T>* tile_gin = reinterpret_cast<Vec4<T>*>(smem_raw);                      // both windows alias the scatter tiles
    Vec4<T>* tile_out = reinterpret_cast<Vec4<T>*>(smem_raw) + kTileNodes;
    int o[3];
    tile_origin(ch, o);
    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) {
        constexpr int n_tiles = ((kMode & BWD_P2G) ? 1 : 0) + ((kMode & BWD_G2P) ? 1 : 0);
        mbar_expect_tx(&mbar, (unsigned)(n_tiles * kTileNodes * sizeof(Vec4<T>)));
        if (kMode & BWD_P2G) tma_load_4d(tile_gin, &tm_gin, &mbar, 0, o[2], o[1], o[0]);
        if (kMode & BWD_G2P) tma_load_4d(tile_out, &tm_out, &mbar, 0, o[2], o[1], o[0]);
    }
    const bool valid = tid < ch.count;
    const int p = ch.start + tid;
    const FramePtr<T> fs = frame_at(frames, s_s.get(), n_pad);
    const FramePtr<T> fprev = frame_at(frames, s_prev.get(), n_pad);
    const FramePtr<T> next = frame_at(adj_next, 0, n_pad);
    const FramePtr<T> cur = frame_at(adj_cur, 0, n_pad);
    V3<T> xs = zero3<T>(), vs = zero3<T>();            // state (x, v) of frame s: p2g.grad input, and G2P(s-1)'s stored output
    V3<T> gx = zero3<T>(), gv = zero3<T>();            // adjoint of (x, v, C) of frame s
    M3<T> gC = zeroM<T>();
    M3<T> C = zeroM<T>(), F = zeroM<T>(), gF_next = zeroM<T>();
    T mu = T(0), lam = T(0), ys = T(0);
    V3<T> part = zero3<T>();
    SvdRec<T> rec;
    if (valid) {
        if (kMode & BWD_P2G) {
            if (kMode & BWD_G2P) prefetch_frame_rest(fprev, p);          // for the next backward kernel (p2g.grad of substep s-1)
            load_xvC(fs, p, xs, vs, C);
            F = load_F(fs, p);
            load_material(P, mat, p, mu, lam, ys);
            const Vec4<T> pa = cur.A0[p];
            part = mk3<T>(pa.x, pa.y, pa.z);
            if (kSvd) rec = load_svd(svd_at(svd_base, s_s.get(), n_pad), p);
            if (!kTwoPhase) gF_next = load_F(next, p);
        } else {
            load_xvC(next, p, gx, gv, gC);
            if (next_ok) {
                const Vec4<T> q0 = fs.A0[p], q1 = fs.A1[p];
                xs = mk3<T>(q0.x, q0.y, q0.z); vs = mk3<T>(q0.w, q1.x, q1.y);
            }
        }
    }
    mbar_wait(&mbar, 0);
    if ((kMode & BWD_P2G) && valid) {
        const Stencil<T> st = make_stencil(xs, P.inv_dx);
        const GridView<T> gview = pick_view(tile_gin, o, g_in, P.n_grid, st.b);
        M3<T> gF;
        p2g_bwd_core<T, kSvd, kTwoPhase>(P, st, gview, vs, C, F, mu, lam, ys, gF_next, part, gx, gv, gC, gF, &rec, &next, p, &fs);
        store_F(cur, p, gF);
        if (!(kMode & BWD_G2P)) store_xvC(cur, p, gx, gv, gC);
    }
    if (kMode & BWD_G2P) {
        G2PBwdCarry<T> carry;
        Stencil<T> stp;
        int key = -1;
        if (valid) {
            const V3<T> xp = load_x(fprev, p);
            stp = make_stencil(xp, P.inv_dx);
            const GridView<T> oview = pick_view(tile_out, o, grid_out, P.n_grid, stp.b);
            V3<T> gxp;
            if ((kMode & BWD_P2G) || next_ok) gxp = g2p_bwd_gather<T, true>(P, xp, stp, oview, gx, gv, gC, xs, vs, carry);
            else gxp = g2p_bwd_gather<T, false>(P, xp, stp, oview, gx, gv, gC, xs, vs, carry);
            ((kMode & BWD_P2G) ? next : cur).A0[p] = mk4<T>(gxp.x, gxp.y, gxp.z, T(0));
            key = pack_cell(stp.b[0], stp.b[1], stp.b[2]);
        }
        __syncthreads();                               // every warp is done with both windows
        Vec4<T>* stile = shared_tile_acquire<T, kBwdTiles>(smem_raw, nullptr, tid >> 5);
        if (valid) {
            WarpTileScatter<T> sc{stile, lane};
            g2p_bwd_scatter<T>(stp, carry, sc);
        }
        flush_runs<T>(stile, lane, key, P.n_grid, g_out, no_peers<Vec4<T>>());
        shared_tile_release<kBwdTiles>(nullptr, tid >> 5, lane);
    }
}

// ------------------------------------------------------------------------------------------------ chunk table
// keys[j] = sorted (block * 64 + cell) of the particle at position j.  The first particle of every block cuts the block's range
// into full chunks of kChunk particles and one remainder (the fewest warps that hold the block's particles) and reserves their
// slots with one atomic.
__device__ __forceinline__ int chunk_origin(int n_grid, unsigned blk) {
    const unsigned nbx = (unsigned)n_grid >> kBlkShift;
    const unsigned bk = blk % nbx, bj = (blk / nbx) % nbx, bi = blk / (nbx * nbx);
    return (int)(((bi << kBlkShift) << 20) | ((bj << kBlkShift) << 10) | (bk << kBlkShift));      // (4 b - 1) + 1 per axis
}
__global__ void k_build_chunks(int n, int n_grid, const unsigned* __restrict__ keys, Chunk* chunks, int* n_chunks, int cap, int* err) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned b = keys[j] >> 6;
    if (j > 0 && (keys[j - 1] >> 6) == b) return;
    int lo = j + 1, hi = n;                              // first index with a larger block id
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((keys[mid] >> 6) > b) hi = mid; else lo = mid + 1;
    }
    const int cnt = lo - j, nch = (cnt + kChunk - 1) / kChunk;
    const int base = atomicAdd(n_chunks, nch);
    if (base + nch > cap) { *err = 3; return; }
    const int origin = chunk_origin(n_grid, b);
    for (int c = 0; c < nch; c++) {
        Chunk ch; ch.origin = origin; ch.start = j + c * kChunk; ch.count = min(kChunk, cnt - c * kChunk); ch.pad = 0;
        chunks[base + c] = ch;
    }
}
// chunk table of an unsorted frame: every chunk claims block 0 (nearly all particles take the dense-grid view); used until the first sort
__global__ void k_trivial_chunks(int n, Chunk* chunks, int* n_chunks) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int nch = (n + kChunk - 1) / kChunk;
    if (c == 0) *n_chunks = nch;
    if (c < nch) { Chunk ch; ch.origin = 0; ch.start = c * kChunk; ch.count = min(kChunk, n - c * kChunk); ch.pad = 0; chunks[c] = ch; }
}
// Window origins from the particles' CURRENT positions (first frame of an env step): origin = (smallest base cell of the chunk's
// particles) - 1 per axis.  Right after the sort that is the block's 4 b - 1; later the material has moved, but the particles
// of a chunk move together, so their base cells still span a few cells and the 8^3 window follows them (spans up to 4 cells
// plus one cell of drift either way within the env step stay inside; what does not fit reads the dense grid, pick_view).
// Without this the windows stay where the blocks were at the sort and a translating body leaves them after a few env steps.
template <class T>
__global__ void __launch_bounds__(kBlock) k_chunk_origins(SimConst<T> P, T* frames, long long n_pad, SlotRef slot, Chunk* chunks, const int* n_chunks) {
    if ((int)blockIdx.x >= *n_chunks) return;
    __shared__ int smin[3][kBlock / 32];
    const Chunk ch = chunks[blockIdx.x];
    int b[3] = {1 << 20, 1 << 20, 1 << 20};
    if ((int)threadIdx.x < ch.count) {
        const V3<T> x = load_x(frame_at(frames, slot.get(), n_pad), ch.start + threadIdx.x);
#pragma unroll
        for (int d = 0; d < 3; d++) b[d] = (int)(x[d] * P.inv_dx - T(0.5));
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b[d] = min(b[d], __shfl_xor_sync(0xffffffffu, b[d], o));
        if ((threadIdx.x & 31) == 0) smin[d][threadIdx.x >> 5] = b[d];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int o[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            int m = smin[d][0];
#pragma unroll
            for (int w = 1; w < kBlock / 32; w++) m = min(m, smin[d][w]);
            o[d] = min(max(m, 0), 1022);          // (origin + 1 is packed into 10 bits per axis; positions are clamped to the domain)
        }
        chunks[blockIdx.x].origin = (o[0] << 20) | (o[1] << 10) | o[2];          // (o - 1) + 1 per axis
    }
}
#endif   // __CUDACC__

}  // namespace plb
