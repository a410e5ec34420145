// __global__ kernels (sm_100a): thin launch wrappers around the per-thread bodies in plb_bodies.cuh, plus the
// reductions (primitive pose gradients, loss terms) that need warp/block cooperation.
#pragma once
#include <cuda_runtime.h>
#include "plb_warp.cuh"

namespace plb {

constexpr int kBlock = 128;
// minimum resident blocks per SM requested from ptxas for the register-heavy adjoint kernels (float only)
template <class T> struct Occ { static constexpr int p2g_bwd = 1, g2p_bwd = 1, g2p_p2g = 1; };
template <> struct Occ<float> { static constexpr int p2g_bwd = 3, g2p_bwd = 3, g2p_p2g = 1; };   // bwd: measured best of {1,3,4} x {1,3,4}

// A frame index given either absolutely (cur == nullptr) or relative to a device-resident cursor.  The cursor form
// lets one captured CUDA graph of an env step (S substeps) be replayed for every env step: only the 3-int cursor
// (slot_in base, slot_out base, primitive-frame base) changes between launches.
struct SlotRef {
    const int* cur; int idx; int rel;
    __device__ __forceinline__ int get() const { return (cur ? cur[idx] : 0) + rel; }
};
__global__ void k_set_cursor(int* cur, int a, int b, int c) { cur[0] = a; cur[1] = b; cur[2] = c; }

// Per-slot compact copy of the forward grid (momentum, mass) of the substep that STARTED at that slot: the list of
// active 4^3 blocks and their 64 Vec4 each.  Written by the forward grid kernel, read by the backward pass instead
// of re-running P2G (SVD + 27-node scatter per particle).
template <class T> struct GridStore {
    Vec4<T>* vals; int* ids; int* cnt; int* overflow; int cap;     // cap blocks per slot; vals == nullptr: disabled
};

// poses of frame pf and pf+1 converted to T in shared memory (2 * n_prim * 8 doubles are read per block)
template <class T>
__device__ __forceinline__ void load_poses_smem(const double* __restrict__ traj, int pf, int n_prim, Pose<T>* s0, Pose<T>* s1) {
    if (threadIdx.x < 2 * n_prim) {
        int which = threadIdx.x >= n_prim ? 1 : 0, k = threadIdx.x - which * n_prim;
        const double* src = traj + ((long long)(pf + which) * PLB_MAX_PRIM + k) * PLB_POSE_DIM;
        Pose<T> p = load_pose<T>(src);
        if (which == 0) s0[k] = p; else s1[k] = p;
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// warp-reduce one PoseGrad and add it to dst[8] (double) with 8 atomics from lane 0
template <class T>
__device__ __forceinline__ void reduce_pose_grad(const PoseGrad<T>& g, double* dst) {
    double vals[8] = {(double)g.pos.x, (double)g.pos.y, (double)g.pos.z, (double)g.rot.w, (double)g.rot.x,
                      (double)g.rot.y, (double)g.rot.z, (double)g.gap};
#pragma unroll
    for (int c = 0; c < 8; c++) {
        double s = warp_sum(vals[c]);
        if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(dst + c, s);
    }
}

// ------------------------------------------------------------------------------------------------ active blocks
// The grid stays dense in memory, but only 4x4x4-node blocks touched by a particle stencil in the current substep are
// visited by the grid kernels.  P2G marks the (up to 8) blocks of each particle in `flags`; k_compact turns the flags
// into a list (and clears them); the grid kernels walk the list.  Invariants: grid_in and g_out are zero outside the
// listed blocks; grid_out / g_in outside the listed blocks are never read.
constexpr int kBlkShift = 2;                 // 4 nodes per block edge
constexpr int kBlkNodes = 64;

__device__ __forceinline__ int block_id(int nbx, int i, int j, int k) {
    return ((i >> kBlkShift) * nbx + (j >> kBlkShift)) * nbx + (k >> kBlkShift);
}

template <class T>
__global__ void k_mark_only(SimConst<T> P, T* frame, long long n_pad, unsigned char* flags) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P.n_particles) mark_blocks<T>(P, load_x(frame_at(frame, 0, n_pad), p), flags);
}

__global__ void k_compact(int n_blocks, unsigned char* flags, int* list, int* count) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_blocks && flags[b]) {
        flags[b] = 0;
        list[atomicAdd(count, 1)] = b;
    }
}

// ---- env-step block list (PLB_ENV_LIST=1): the active-block list is built ONCE per env step from the frame the step starts
// at, dilated by one block in every direction (a particle moves far less than 4 nodes in one env step), and every substep
// of the step walks that fixed list: no per-substep flag marking, memset or compaction.  k_check_listed verifies at the
// end of the step that the frame it produced still lies inside the list.
template <class T>
__global__ void k_mark_slot(SimConst<T> P, T* frames, long long n_pad, SlotRef slot, unsigned char* flags) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P.n_particles) mark_blocks<T>(P, load_x(frame_at(frames, slot.get(), n_pad), p), flags);
}
// out[b'] = 1 for the 27 neighbours b' of every flagged block b; clears in[b]
// slab runs: every stencil of this rank's particles must stay inside [lo, hi) planes (owned planes + halo); *err = 2 otherwise
template <class T>
__global__ void k_check_margin(SimConst<T> P, T* frames, long long n_pad, SlotRef slot, int lo, int hi, int* err) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    const V3<T> x = load_x(frame_at(frames, slot.get(), n_pad), p);
    const int b = (int)(x.x * P.inv_dx - T(0.5));
    if (b < lo || b + 2 >= hi) *err = 2;
}
__global__ void k_dilate_flags(int nbx, unsigned char* in, unsigned char* out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbx * nbx * nbx || !in[b]) return;
    in[b] = 0;
    const int bk = b % nbx, bj = (b / nbx) % nbx, bi = b / (nbx * nbx);
    for (int di = -1; di <= 1; di++)
        for (int dj = -1; dj <= 1; dj++)
            for (int dk = -1; dk <= 1; dk++) {
                const int i = bi + di, j = bj + dj, k = bk + dk;
                if (i >= 0 && j >= 0 && k >= 0 && i < nbx && j < nbx && k < nbx) out[(i * nbx + j) * nbx + k] = 1;
            }
}
// compaction that also records membership (listed[b] = 1 / 0 for every block) and clears the flags
__global__ void k_compact_mark(int n_blocks, unsigned char* flags, int* list, int* count, unsigned char* listed) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const bool on = flags[b] != 0;
    listed[b] = on ? 1 : 0;
    if (on) {
        flags[b] = 0;
        list[atomicAdd(count, 1)] = b;
    }
}
// every flagged block must be listed; clears the flags; *err = 2 otherwise
__global__ void k_check_listed(int n_blocks, unsigned char* flags, const unsigned char* __restrict__ listed, int* err) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks || !flags[b]) return;
    flags[b] = 0;
    if (!listed[b]) *err = 2;
}

// node handled by thread `local` (0..63) of listed block `blk`
__device__ __forceinline__ long long block_node(int n_grid, int blk, int local) {
    const int nbx = n_grid >> kBlkShift;
    int bk = blk % nbx, bj = (blk / nbx) % nbx, bi = blk / (nbx * nbx);
    int i = (bi << kBlkShift) + (local >> 4), j = (bj << kBlkShift) + ((local >> 2) & 3), k = (bk << kBlkShift) + (local & 3);
    return ((long long)i * n_grid + j) * n_grid + k;
}

// ------------------------------------------------------------------------------------------------ slab halo
// Multi-GPU slab decomposition along grid axis 0 (planes are contiguous: zero-copy sends).  A zone is the plane range
// [lo, hi) around a slab boundary; `recv` holds the neighbour's partial sums for exactly those planes.
//  k_halo_mark: blocks of the zone that lie in MY owned planes and carry mass from the neighbour become active here too
//  (the owner of a plane is responsible for its nodes' pose gradients).
//  k_halo_add_listed: add the neighbour's values into my listed blocks that lie inside the zone.
template <class T>
__global__ void k_halo_mark(int n_grid, const Vec4<T>* __restrict__ recv, int zone_lo, int zone_hi, int own_lo, int own_hi,
                            unsigned char* flags) {
    const int nbx = n_grid >> kBlkShift;
    const int lo = max(zone_lo, own_lo), hi = min(zone_hi, own_hi);
    const int nb_planes = (hi - lo) >> kBlkShift;
    const int per_cta = blockDim.x / kBlkNodes, local = threadIdx.x & (kBlkNodes - 1);
    const int total = nb_planes * nbx * nbx;
    for (int e = blockIdx.x * per_cta + threadIdx.x / kBlkNodes; e < total; e += gridDim.x * per_cta) {
        int bi = (lo >> kBlkShift) + e / (nbx * nbx), bj = (e / nbx) % nbx, bk = e % nbx;
        int i = (bi << kBlkShift) + (local >> 4), j = (bj << kBlkShift) + ((local >> 2) & 3), k = (bk << kBlkShift) + (local & 3);
        Vec4<T> v = recv[((long long)(i - zone_lo) * n_grid + j) * n_grid + k];
        if (v.w > T(0)) flags[(bi * nbx + bj) * nbx + bk] = 1;
    }
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_halo_add_listed(int n_grid, Vec4<T>* grid, const Vec4<T>* __restrict__ recv, int zone_lo,
                                                            int zone_hi, const int* __restrict__ list, const int* __restrict__ count) {
    const int per_cta = kBlock / kBlkNodes, local = threadIdx.x & (kBlkNodes - 1), n = *count;
    const int nbx = n_grid >> kBlkShift;
    for (int e = blockIdx.x * per_cta + threadIdx.x / kBlkNodes; e < n; e += gridDim.x * per_cta) {
        const int blk = list[e];
        const int i0 = (blk / (nbx * nbx)) << kBlkShift;
        if (i0 < zone_lo || i0 + 4 > zone_hi) continue;
        const long long node = block_node(n_grid, blk, local);
        const long long plane_sz = (long long)n_grid * n_grid;
        Vec4<T> r = recv[node - (long long)zone_lo * plane_sz];
        Vec4<T> g = grid[node];
        grid[node] = mk4<T>(g.x + r.x, g.y + r.y, g.z + r.z, g.w + r.w);
    }
}

// ---- peer-memory halo (NVLink P2P stores into the neighbour's inbox, no host round trip) -------------------------------
// Inbox of one side (lives in the RECEIVER's memory, mapped into the sender through CUDA IPC):
//   int flag[64]                      flag[0] = sequence number of the newest complete push
//   int stamps[2][zone_blocks]        per parity: sequence number at which the block's data were pushed
//   Vec4 data[2][zone_blocks][64]     per parity: the block's 64 node values (block-major, same order as the store)
// A push writes only the sender's ACTIVE blocks inside the zone; the receiver trusts a block iff its stamp == seq.
//   uchar zflags[2][zone_blocks]      per parity: the sender's block flags of the zone (env-step list exchange)
struct HaloGeom { int zone_lo, zone_hi, nzb; long long stamps_off, data_off, flags_off; };   // offsets in bytes from the inbox base

// What a grid kernel needs to take the neighbours' partial sums itself (fused receive): my two inboxes, their geometry, the
// exchange counter.  on == 0: no halo (single GPU, or the stage reads an already summed grid).
// on == 3: the grid kernel also SENDS (halo_push_share): peer = the neighbours' inboxes (mapped through CUDA IPC), pg = their geometry,
// seq_w / done = the exchange counter and the CTA counter of the last-CTA publication.
struct HaloIn { const char* inbox[2]; HaloGeom g[2]; const int* seq; int* err; int on; char* peer[2]; HaloGeom pg[2]; int* seq_w; unsigned* done; };
__host__ __device__ __forceinline__ HaloIn halo_none() {
    HaloIn h; h.inbox[0] = h.inbox[1] = nullptr; h.seq = nullptr; h.err = nullptr; h.on = 0; h.peer[0] = h.peer[1] = nullptr; h.seq_w = nullptr; h.done = nullptr; return h;
}

// thread 0 of the CTA spins until both neighbours have published exchange `s` in MY inboxes (local memory: the neighbour wrote
// it over NVLink), then the CTA may read what they pushed.  ~4 s timeout -> *err = 1 (reported by the next readback).
__device__ __forceinline__ int halo_wait_cta(const HaloIn& h, int s_given = -1) {
    const int s = s_given >= 0 ? s_given : *h.seq;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        const volatile int* e = h.err;
        for (int side = 0; side < 2; side++) {
            if (!h.inbox[side]) continue;
            const volatile int* f = reinterpret_cast<const volatile int*>(h.inbox[side]);
            while (f[0] < s) {
                if (*e == 1) break;                   // an earlier wait of this run timed out: do not spin again (the run is reported invalid)
                if (clock64() - t0 > 8000000000LL) { *h.err = 1; break; }
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    return s;
}
// the neighbours' contribution to node `local` of listed block `blk` in exchange s (zero if the block is outside the zones or
// was not pushed); inbox data bypass L1 (the same addresses held the exchange before last)
template <class T>
__device__ __forceinline__ bool halo_fetch(const HaloIn& h, int n_grid, int blk, int local, int s, Vec4<T>& out) {
    const int nbx = n_grid >> kBlkShift;
    const int i0 = (blk / (nbx * nbx)) << kBlkShift;
    bool any = false;
    out = mk4<T>(T(0), T(0), T(0), T(0));
#pragma unroll
    for (int side = 0; side < 2; side++) {
        if (!h.inbox[side]) continue;
        const HaloGeom& g = h.g[side];
        if (i0 < g.zone_lo || i0 + 4 > g.zone_hi) continue;
        const int zb = (i0 - g.zone_lo) / 4 * nbx * nbx + blk % (nbx * nbx);
        const int par = s & 1;
        const int* stamps = reinterpret_cast<const int*>(h.inbox[side] + g.stamps_off) + (long long)par * g.nzb;
        if (__ldcg(stamps + zb) != s) continue;
        const Vec4<T>* data = reinterpret_cast<const Vec4<T>*>(h.inbox[side] + g.data_off) + ((long long)par * g.nzb + zb) * kBlkNodes;
        const T* q = reinterpret_cast<const T*>(data + local);
        out.x += __ldcg(q); out.y += __ldcg(q + 1); out.z += __ldcg(q + 2); out.w += __ldcg(q + 3);
        any = true;
    }
    return any;
}

__device__ __forceinline__ int zone_block_index(int n_grid, int blk, int zone_lo) {
    const int nbx = n_grid >> kBlkShift;
    const int bi = blk / (nbx * nbx);
    return (bi - (zone_lo >> kBlkShift)) * nbx * nbx + blk % (nbx * nbx);
}

// sender: copy my listed blocks that lie in the zone into the peer's inbox (parity = seq & 1) and stamp them
template <class T>
__global__ void __launch_bounds__(kBlock) k_halo_push(int n_grid, const Vec4<T>* __restrict__ grid, const int* __restrict__ list,
                                                      const int* __restrict__ count, char* peer_inbox, HaloGeom g, const int* seq_ptr) {
    const int per_cta = kBlock / kBlkNodes, local = threadIdx.x & (kBlkNodes - 1), n = *count, seq = *seq_ptr, par = seq & 1;
    const int nbx = n_grid >> kBlkShift;
    int* stamps = reinterpret_cast<int*>(peer_inbox + g.stamps_off) + (long long)par * g.nzb;
    Vec4<T>* data = reinterpret_cast<Vec4<T>*>(peer_inbox + g.data_off) + (long long)par * g.nzb * kBlkNodes;
    for (int e = blockIdx.x * per_cta + threadIdx.x / kBlkNodes; e < n; e += gridDim.x * per_cta) {
        const int blk = list[e];
        const int i0 = (blk / (nbx * nbx)) << kBlkShift;
        if (i0 < g.zone_lo || i0 + 4 > g.zone_hi) continue;
        const int zb = zone_block_index(n_grid, blk, g.zone_lo);
        data[(long long)zb * kBlkNodes + local] = grid[block_node(n_grid, blk, local)];
        if (local == 0) stamps[zb] = seq;
    }
}
// one thread: make the pushes visible system-wide, then publish the sequence number in the peers' flags
__global__ void k_halo_signal(char* peer0, char* peer1, const int* seq_ptr) {
    __threadfence_system();
    const int seq = *seq_ptr;
    if (peer0) { volatile int* f = reinterpret_cast<volatile int*>(peer0); f[0] = seq; }
    if (peer1) { volatile int* f = reinterpret_cast<volatile int*>(peer1); f[0] = seq; }
    __threadfence_system();
}
// one thread: spin until both neighbours published `seq` in MY inboxes; err is set on a ~4 s timeout
__global__ void k_halo_wait(const char* inbox0, const char* inbox1, const int* seq_ptr, int* err) {
    const int seq = *seq_ptr;
    const long long t0 = clock64();
    for (int side = 0; side < 2; side++) {
        const char* ib = side == 0 ? inbox0 : inbox1;
        if (!ib) continue;
        const volatile int* f = reinterpret_cast<const volatile int*>(ib);
        while (f[0] < seq) {
            if (clock64() - t0 > 8000000000LL) { *err = 1; break; }
        }
    }
    __threadfence_system();
}
__global__ void k_halo_next_seq(int* seq_ptr) { *seq_ptr += 1; }

// ---- fused halo (slab runs with the env-step block list): ONE launch per exchange on the sending side, none on the receiving
// side (the grid kernel that consumes the data waits for it itself, halo_wait_cta / halo_fetch).
// Exchange number s = *seq + 1.  Every CTA copies its share of my listed zone blocks into the neighbours' inboxes (parity
// s & 1) and stamps them; every thread fences its stores; the LAST CTA to finish publishes s in both neighbours' flags and
// stores it to *seq (the consumer kernel launched next reads it there).
__device__ __forceinline__ void halo_publish_last_cta(char* peer0, char* peer1, int* seq_ptr, unsigned* done, int s, bool fence = true) {
    if (fence) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(done, 1u);
        if (t == gridDim.x - 1) {
            *done = 0u;
            __threadfence_system();
            if (s < 0) s = *seq_ptr + 1;
            if (peer0) { volatile int* f = reinterpret_cast<volatile int*>(peer0); f[0] = s; }
            if (peer1) { volatile int* f = reinterpret_cast<volatile int*>(peer1); f[0] = s; }
            __threadfence_system();
            *seq_ptr = s;
        }
    }
}
// Direct halo: the scatter kernels add their zone contributions straight into the neighbours' grids (PeerHalo, plb_warp.cuh), so
// an exchange is only a completion signal: every warp that issued remote REDs fences them, the last CTA of the scatter kernel
// publishes the next exchange number in the neighbours' flags.  seq == nullptr: nothing to publish (single GPU).
struct HaloOut { char* peer[2]; int* seq; unsigned* done; };
__host__ __device__ __forceinline__ HaloOut halo_out_none() { HaloOut h; h.peer[0] = h.peer[1] = nullptr; h.seq = nullptr; h.done = nullptr; return h; }
// every thread of the CTA calls this once, at the end of the kernel (sent: this thread issued a RED into a neighbour's grid)
__device__ __forceinline__ void halo_publish_scatter(const HaloOut& ho, bool sent) {
    if (!ho.seq) return;
    if (__any_sync(0xffffffffu, sent)) __threadfence_system();
    halo_publish_last_cta(ho.peer[0], ho.peer[1], ho.seq, ho.done, -1, false);
}
template <class T>
__global__ void __launch_bounds__(kBlock) k_halo_push2(int n_grid, const Vec4<T>* __restrict__ grid, const int* __restrict__ list,
                                                       const int* __restrict__ count, char* peer0, char* peer1, HaloGeom g0, HaloGeom g1,
                                                       int* seq_ptr, unsigned* done) {
    const int per_cta = kBlock / kBlkNodes, local = threadIdx.x & (kBlkNodes - 1), n = *count, s = *seq_ptr + 1, par = s & 1;
    const int nbx = n_grid >> kBlkShift;
    for (int e = blockIdx.x * per_cta + threadIdx.x / kBlkNodes; e < n; e += gridDim.x * per_cta) {
        const int blk = list[e];
        const int i0 = (blk / (nbx * nbx)) << kBlkShift;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            char* peer = side == 0 ? peer0 : peer1;
            const HaloGeom& g = side == 0 ? g0 : g1;
            if (!peer || i0 < g.zone_lo || i0 + 4 > g.zone_hi) continue;
            const int zb = zone_block_index(n_grid, blk, g.zone_lo);
            int* stamps = reinterpret_cast<int*>(peer + g.stamps_off) + (long long)par * g.nzb;
            Vec4<T>* data = reinterpret_cast<Vec4<T>*>(peer + g.data_off) + (long long)par * g.nzb * kBlkNodes;
            data[(long long)zb * kBlkNodes + local] = grid[block_node(n_grid, blk, local)];
            if (local == 0) stamps[zb] = s;
        }
    }
    halo_publish_last_cta(peer0, peer1, seq_ptr, done, s);
}
// The same push from inside the consuming grid kernel (HaloIn.on == 3): every CTA copies its share of my listed zone blocks of
// `grid` into the neighbours' inboxes, the last CTA to finish publishes exchange s = *seq + 1 (every CTA read *seq before it
// counted itself done, so the publisher's update of *seq cannot be seen early).  One launch less per exchange than k_halo_push2 +
// grid kernel, and the kernel runs its interior blocks while the neighbours' pushes are in flight.
__device__ __forceinline__ bool halo_zone_block(const HaloIn& h, int n_grid, int blk) {
    const int nbx = n_grid >> kBlkShift;
    const int i0 = (blk / (nbx * nbx)) << kBlkShift;
    return (h.inbox[0] && i0 >= h.g[0].zone_lo && i0 + 4 <= h.g[0].zone_hi) || (h.inbox[1] && i0 >= h.g[1].zone_lo && i0 + 4 <= h.g[1].zone_hi);
}
template <class T>
__device__ __forceinline__ int halo_push_share(const HaloIn& h, int n_grid, const Vec4<T>* grid, const int* __restrict__ list, int n) {
    const int per_cta = kBlock / kBlkNodes, local = threadIdx.x & (kBlkNodes - 1), s = *h.seq + 1, par = s & 1;
    const int nbx = n_grid >> kBlkShift;
    for (int e = blockIdx.x * per_cta + threadIdx.x / kBlkNodes; e < n; e += gridDim.x * per_cta) {
        const int blk = list[e];
        const int i0 = (blk / (nbx * nbx)) << kBlkShift;
#pragma unroll
        for (int side = 0; side < 2; side++) {
            char* peer = h.peer[side];
            const HaloGeom& g = h.pg[side];
            if (!peer || i0 < g.zone_lo || i0 + 4 > g.zone_hi) continue;
            const int zb = zone_block_index(n_grid, blk, g.zone_lo);
            int* stamps = reinterpret_cast<int*>(peer + g.stamps_off) + (long long)par * g.nzb;
            Vec4<T>* data = reinterpret_cast<Vec4<T>*>(peer + g.data_off) + (long long)par * g.nzb * kBlkNodes;
            data[(long long)zb * kBlkNodes + local] = grid[block_node(n_grid, blk, local)];
            if (local == 0) stamps[zb] = s;
        }
    }
    halo_publish_last_cta(h.peer[0], h.peer[1], h.seq_w, h.done, s);
    return s;
}
// env-step list exchange: my block flags inside the zones -> the neighbours' inboxes (one byte per zone block), published like a push
__global__ void k_halo_push_flags(int n_grid, const unsigned char* __restrict__ flags, char* peer0, char* peer1, HaloGeom g0, HaloGeom g1,
                                  int* seq_ptr, unsigned* done) {
    const int s = *seq_ptr + 1, par = s & 1, nbx = n_grid >> kBlkShift;
    for (int side = 0; side < 2; side++) {
        char* peer = side == 0 ? peer0 : peer1;
        const HaloGeom& g = side == 0 ? g0 : g1;
        if (!peer) continue;
        unsigned char* dst = reinterpret_cast<unsigned char*>(peer + g.flags_off) + (long long)par * g.nzb;
        const int first = (g.zone_lo >> kBlkShift) * nbx * nbx;
        for (int zb = blockIdx.x * blockDim.x + threadIdx.x; zb < g.nzb; zb += gridDim.x * blockDim.x) dst[zb] = flags[first + zb];
    }
    halo_publish_last_cta(peer0, peer1, seq_ptr, done, s);
}
// ... and OR what the neighbours sent into my flags (whole zones: the list is dilated afterwards, also across the ownership boundary)
__global__ void k_halo_or_flags(int n_grid, unsigned char* flags, HaloIn h) {
    const int s = halo_wait_cta(h), par = s & 1, nbx = n_grid >> kBlkShift;
    for (int side = 0; side < 2; side++) {
        if (!h.inbox[side]) continue;
        const HaloGeom& g = h.g[side];
        const unsigned char* src = reinterpret_cast<const unsigned char*>(h.inbox[side] + g.flags_off) + (long long)par * g.nzb;
        const int first = (g.zone_lo >> kBlkShift) * nbx * nbx;
        for (int zb = blockIdx.x * blockDim.x + threadIdx.x; zb < g.nzb; zb += gridDim.x * blockDim.x)
            if (__ldcg(src + zb)) flags[first + zb] = 1;
    }
}

// receiver, forward only: blocks of the zone in MY owned planes that the neighbour pushed but I do not list yet are
// appended to my list (the owner of a plane accumulates its nodes' pose gradients)
__global__ void k_halo_append(int n_grid, const char* inbox, HaloGeom g, int own_lo, int own_hi, const int* seq_ptr, int* listed_stamp,
                              int* list, int* count) {
    const int seq = *seq_ptr, par = seq & 1, nbx = n_grid >> kBlkShift;
    const int* stamps = reinterpret_cast<const int*>(inbox + g.stamps_off) + (long long)par * g.nzb;
    for (int zb = blockIdx.x * blockDim.x + threadIdx.x; zb < g.nzb; zb += gridDim.x * blockDim.x) {
        if (stamps[zb] != seq) continue;
        const int bi = (g.zone_lo >> kBlkShift) + zb / (nbx * nbx);
        const int i0 = bi << kBlkShift;
        if (i0 < own_lo || i0 >= own_hi) continue;
        const int blk = bi * nbx * nbx + zb % (nbx * nbx);
        if (listed_stamp[blk] == seq) continue;
        listed_stamp[blk] = seq;
        list[atomicAdd(count, 1)] = blk;
    }
}
// receiver: add the neighbour's values of this sequence number into my listed blocks
template <class T>
__global__ void __launch_bounds__(kBlock) k_halo_add_inbox(int n_grid, Vec4<T>* grid, const char* inbox, HaloGeom g,
                                                           const int* __restrict__ list, const int* __restrict__ count, const int* seq_ptr) {
    const int per_cta = kBlock / kBlkNodes, local = threadIdx.x & (kBlkNodes - 1), n = *count, seq = *seq_ptr, par = seq & 1;
    const int nbx = n_grid >> kBlkShift;
    const int* stamps = reinterpret_cast<const int*>(inbox + g.stamps_off) + (long long)par * g.nzb;
    const Vec4<T>* data = reinterpret_cast<const Vec4<T>*>(inbox + g.data_off) + (long long)par * g.nzb * kBlkNodes;
    for (int e = blockIdx.x * per_cta + threadIdx.x / kBlkNodes; e < n; e += gridDim.x * per_cta) {
        const int blk = list[e];
        const int i0 = (blk / (nbx * nbx)) << kBlkShift;
        if (i0 < g.zone_lo || i0 + 4 > g.zone_hi) continue;
        const int zb = zone_block_index(n_grid, blk, g.zone_lo);
        if (stamps[zb] != seq) continue;
        const long long node = block_node(n_grid, blk, local);
        Vec4<T> r = data[(long long)zb * kBlkNodes + local];
        Vec4<T> v = grid[node];
        grid[node] = mk4<T>(v.x + r.x, v.y + r.y, v.z + r.z, v.w + r.w);
    }
}
// compaction that also records which blocks are listed at this sequence number (slab peer mode)
__global__ void k_compact_stamped(int n_blocks, unsigned char* flags, int* list, int* count, int* listed_stamp, const int* seq_ptr) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_blocks && flags[b]) {
        flags[b] = 0;
        listed_stamp[b] = *seq_ptr;
        list[atomicAdd(count, 1)] = b;
    }
}
__global__ void k_stamp_list(const int* __restrict__ list, const int* __restrict__ count, int* listed_stamp, const int* seq_ptr) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < *count) listed_stamp[list[e]] = *seq_ptr;
}

template <class T> __global__ void k_add_scalar(long long n, T* dst, const T* __restrict__ src) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// ------------------------------------------------------------------------------------------------ substep
template <class T>
__global__ void __launch_bounds__(kBlock) k_p2g(SimConst<T> P, T* frames, long long n_pad, SlotRef slot_in, SlotRef slot_out,
                                                int store_F_out, Material<T> mat, Vec4<T>* grid_in, unsigned char* flags, T* svd_base) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    FramePtr<T> fin = frame_at(frames, slot_in.get(), n_pad);
    SvdPtr<T> sp = svd_at(svd_base, slot_in.get(), n_pad);
    p2g_body<T>(p, P, fin, frame_at(frames, slot_out.get(), n_pad), store_F_out != 0, mat, grid_in, svd_base ? &sp : nullptr);
    if (flags) mark_blocks<T>(P, load_x(fin, p), flags);
}

// ---- scatter kernels with the warp tile (plb_warp.cuh); dynamic shared memory = (blockDim / 32) tiles of 14.25 KB.
template <class T> __device__ __forceinline__ Vec4<T>* warp_tile_ptr(unsigned char* smem_raw) {
    return reinterpret_cast<Vec4<T>*>(smem_raw) + (threadIdx.x >> 5) * kTileVec4;
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_p2g_warp(SimConst<T> P, T* frames, long long n_pad, SlotRef slot_in, SlotRef slot_out,
                                                     int store_F_out, Material<T> mat, Vec4<T>* grid_in, unsigned char* flags, int flush_mode, T* svd_base) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SvdPtr<T> sp = svd_at(svd_base, slot_in.get(), n_pad);
    t_p2g<T>(blockIdx.x * blockDim.x + threadIdx.x, threadIdx.x & 31, warp_tile_ptr<T>(smem_raw), P,
                     frame_at(frames, slot_in.get(), n_pad), frame_at(frames, slot_out.get(), n_pad), store_F_out != 0, mat, grid_in, flags, flush_mode,
                     svd_base ? &sp : nullptr);
}

// G2P of substep s + P2G of substep s+1 in one pass over the particles (inside env-step graphs)
// kMinB: resident CTAs per SM requested from ptxas (register cap 65536 / (128 kMinB)): 5 -> 96 registers (what ptxas picks
// unprompted), 6 -> 80 registers
template <class T, int kMinB>
__global__ void __launch_bounds__(kBlock, kMinB) k_g2p_p2g_warp(SimConst<T> P, T* frames, long long n_pad, SlotRef slot_in,
                                                                           SlotRef slot_mid, SlotRef slot_out, Material<T> mat,
                                                                           const Vec4<T>* grid_out, Vec4<T>* grid_in, unsigned char* flags, int flush_mode, T* svd_base) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SvdPtr<T> sp = svd_at(svd_base, slot_mid.get(), n_pad);          // the P2G half decomposes F_tmp of frame `mid`
    t_g2p_p2g<T>(blockIdx.x * blockDim.x + threadIdx.x, threadIdx.x & 31, warp_tile_ptr<T>(smem_raw), P,
                         frame_at(frames, slot_in.get(), n_pad), frame_at(frames, slot_mid.get(), n_pad),
                         frame_at(frames, slot_out.get(), n_pad), mat, grid_out, grid_in, flags, flush_mode, svd_base ? &sp : nullptr);
}

// g2p.grad; next_ok: slot_in + 1 holds the frame G2P produced from slot_in (clamp masks and gather sum are read from it)
template <class T>
__global__ void __launch_bounds__(kBlock, Occ<T>::g2p_bwd) k_g2p_bwd_warp(SimConst<T> P, T* frames, long long n_pad, SlotRef slot_in, int next_ok,
                                                                           T* adj_next, T* adj_cur, const Vec4<T>* grid_out, Vec4<T>* g_out, int flush_mode,
                                                                           PeerHalo<Vec4<T>> ph, HaloOut ho) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int si = slot_in.get();
    FramePtr<T> fnext = frame_at(frames, si + 1, n_pad);
    const bool sent = t_g2p_bwd<T>(blockIdx.x * blockDim.x + threadIdx.x, threadIdx.x & 31, warp_tile_ptr<T>(smem_raw), P,
                         frame_at(frames, si, n_pad), next_ok ? &fnext : nullptr, frame_at(adj_next, 0, n_pad), frame_at(adj_cur, 0, n_pad),
                         grid_out, g_out, flush_mode, ph.any() ? &ph : nullptr);
    halo_publish_scatter(ho, sent);
}

// p2g.grad of substep s + g2p.grad of substep s-1 (inside env-step graphs; frame s was produced by G2P(s-1) there)
// kSvd: the decomposition of F_tmp(s) comes from the SVD store (written by the forward pass) instead of the Jacobi iteration
template <class T, int kMinB, bool kSvd>
__global__ void __launch_bounds__(kBlock, kMinB) k_p2g_bwd_g2p_bwd_warp(SimConst<T> P, T* frames, long long n_pad, SlotRef slot_s,
                                                                                   SlotRef slot_prev, T* adj_next, T* adj_cur, Material<T> mat,
                                                                                   const Vec4<T>* g_in, const Vec4<T>* grid_out, Vec4<T>* g_out, int flush_mode, T* svd_base,
                                                                                   PeerHalo<Vec4<T>> ph, HaloOut ho) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pdl_launch();                                              // (the grid adjoint behind this kernel waits for its completion itself)
    SvdPtr<T> sp = svd_at(svd_base, slot_s.get(), n_pad);
    // tight register cap + SVD store: run the (then cheap) forward particle math twice instead of keeping it across the gather
    const bool sent = t_p2g_bwd_g2p_bwd<T, kSvd, (kSvd && kMinB >= 4)>(blockIdx.x * blockDim.x + threadIdx.x, threadIdx.x & 31, warp_tile_ptr<T>(smem_raw), P,
                                       frame_at(frames, slot_s.get(), n_pad), frame_at(frames, slot_prev.get(), n_pad), frame_at(adj_next, 0, n_pad),
                                       frame_at(adj_cur, 0, n_pad), mat, g_in, grid_out, g_out, flush_mode, &sp, ph.any() ? &ph : nullptr);
    halo_publish_scatter(ho, sent);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_grid_fwd_sparse(SimConst<T> P, PrimSet<T> prims, const double* traj, SlotRef pf,
                                                            Vec4<T>* grid_in, Vec4<T>* grid_out, int clear_in,
                                                            const int* __restrict__ list, const int* __restrict__ count,
                                                            GridStore<T> store, SlotRef slot, HaloIn halo) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    load_poses_smem<T>(traj, pf.get(), P.n_prim, s0, s1);
    const int per_cta = kBlock / kBlkNodes, n = *count;
    const int local = threadIdx.x & (kBlkNodes - 1);
    pdl_wait();                                                // the scatter into grid_in is complete
    pdl_launch();                                              // the next particle kernel may load its particles meanwhile
    int hs = 0;
    if (halo.on == 3) hs = halo_push_share<T>(halo, P.n_grid, grid_in, list, n);      // send first; the wait comes after the interior blocks
    else if (halo.on) hs = halo_wait_cta(halo);                // fused receive: the neighbours' partial sums of this exchange
    Vec4<T>* svals = nullptr; int* sids = nullptr;
    if (store.vals) {
        const long long sl = slot.get();
        svals = store.vals + sl * store.cap * kBlkNodes;
        sids = store.ids + sl * store.cap;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            store.cnt[sl] = n <= store.cap ? n : -1;
            if (n > store.cap) *store.overflow = 1;
        }
    }
    // on == 3: pass 0 = blocks outside the zones (while the neighbours' pushes arrive), then wait, pass 1 = zone blocks
    for (int pass = 0; pass < (halo.on == 3 ? 2 : 1); pass++) {
    if (pass == 1) halo_wait_cta(halo, hs);
    for (int e = blockIdx.x * per_cta + threadIdx.x / kBlkNodes; e < n; e += gridDim.x * per_cta) {
        const int blk = list[e];
        if (halo.on == 3 && halo_zone_block(halo, P.n_grid, blk) != (pass == 1)) continue;
        const long long node = block_node(P.n_grid, blk, local);
        if (halo.on == 1 || (halo.on == 3 && pass == 1)) {      // (2: direct halo, the neighbours' sums are already in grid_in)
            Vec4<T> r;
            if (halo_fetch<T>(halo, P.n_grid, blk, local, hs, r)) {
                const Vec4<T> v = grid_in[node];
                grid_in[node] = mk4<T>(v.x + r.x, v.y + r.y, v.z + r.z, v.w + r.w);
            }
        }
        if (svals && e < store.cap) {
            svals[(long long)e * kBlkNodes + local] = grid_in[node];
            if (local == 0) sids[e] = blk;
        }
        grid_fwd_body<T>(node, P, prims, s0, s1, grid_in, grid_out, clear_in != 0);
    }
    }
}

// backward: re-install the stored forward grid of `slot` (values + active list) instead of recomputing P2G
template <class T>
__global__ void __launch_bounds__(kBlock) k_restore_blocks(int n_grid, Vec4<T>* grid_in, int* list, int* count, GridStore<T> store, SlotRef slot) {
    const long long sl = slot.get();
    const int n = min(store.cnt[sl], store.cap);          // (-1 or > cap: overflow, flagged by the forward kernel)
    const Vec4<T>* svals = store.vals + sl * store.cap * kBlkNodes;
    const int* sids = store.ids + sl * store.cap;
    const int per_cta = kBlock / kBlkNodes, local = threadIdx.x & (kBlkNodes - 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) *count = n;
    for (int e = blockIdx.x * per_cta + threadIdx.x / kBlkNodes; e < n; e += gridDim.x * per_cta) {
        const int blk = sids[e];
        grid_in[block_node(n_grid, blk, local)] = svals[(long long)e * kBlkNodes + local];
        if (local == 0) list[e] = blk;
    }
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_grid_bwd_sparse(SimConst<T> P, PrimSet<T> prims, const double* traj, SlotRef pfr,
                                                            Vec4<T>* grid_in, Vec4<T>* g_out, Vec4<T>* g_in, int clear,
                                                            double* prim_grad, const int* __restrict__ list,
                                                            const int* __restrict__ count, int own_lo, int own_hi) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    const int pf = pfr.get();
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    const int per_cta = kBlock / kBlkNodes, n = *count;
    const int rounds = (n + gridDim.x * per_cta - 1) / (gridDim.x * per_cta);
    for (int r = 0; r < rounds; r++) {              // uniform trip count: the pose-gradient reduction uses warp shuffles
        int e = (r * gridDim.x + blockIdx.x) * per_cta + threadIdx.x / kBlkNodes;
        PoseGrad<T> g0[PLB_MAX_PRIM], g1[PLB_MAX_PRIM];
        unsigned touched = 0;
        if (e < n) {
            for (int k = 0; k < P.n_prim; k++) { g0[k].clear(); g1[k].clear(); }
            const long long node = block_node(P.n_grid, list[e], threadIdx.x & (kBlkNodes - 1));
            grid_bwd_body<T>(node, P, prims, s0, s1, grid_in, g_out, g_in, clear != 0, g0, g1, touched);
            // slab decomposition: a node's pose gradient is accumulated by the rank that owns its plane
            const int plane = (int)(node / ((long long)P.n_grid * P.n_grid));
            if (plane < own_lo || plane >= own_hi) touched = 0;
        }
        for (int k = 0; k < P.n_prim; k++) {
            unsigned any = __ballot_sync(0xffffffffu, (touched >> k) & 1u);
            if (any == 0) continue;
            if (!((touched >> k) & 1u)) { g0[k].clear(); g1[k].clear(); }
            reduce_pose_grad<T>(g0[k], prim_grad + ((long long)pf * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
            reduce_pose_grad<T>(g1[k], prim_grad + ((long long)(pf + 1) * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
        }
    }
}

// same operator, pose gradients of one primitive at a time in registers (t_grid_bwd_node, plb_warp.cuh): no per-thread
// PoseGrad arrays in local memory, float warp reduction, one pass
template <class T>
__global__ void __launch_bounds__(kBlock) k_grid_bwd_sparse_v2(SimConst<T> P, PrimSet<T> prims, const double* traj, SlotRef pfr,
                                                               Vec4<T>* grid_in, Vec4<T>* g_out, Vec4<T>* g_in, int clear,
                                                               double* prim_grad, const int* __restrict__ list,
                                                               const int* __restrict__ count, int own_lo, int own_hi, HaloIn halo) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    const int pf = pfr.get();
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    const int per_cta = kBlock / kBlkNodes, n = *count;
    const int rounds = (n + gridDim.x * per_cta - 1) / (gridDim.x * per_cta);
    pdl_wait();                                                // the scatter into g_out is complete
    pdl_launch();
    int hs = 0;
    if (halo.on == 3) hs = halo_push_share<T>(halo, P.n_grid, g_out, list, n);
    else if (halo.on) hs = halo_wait_cta(halo);                // fused receive of the neighbours' adjoint of grid_out
    for (int pass = 0; pass < (halo.on == 3 ? 2 : 1); pass++) {
    if (pass == 1) halo_wait_cta(halo, hs);
    for (int r = 0; r < rounds; r++) {              // uniform trip count: warp collectives inside
        const int e = (r * gridDim.x + blockIdx.x) * per_cta + threadIdx.x / kBlkNodes;
        bool act = e < n;
        long long node = 0;
        bool owned = false;
        if (act && halo.on == 3 && halo_zone_block(halo, P.n_grid, list[e]) != (pass == 1)) act = false;
        if (act) {
            const int blk = list[e], local = threadIdx.x & (kBlkNodes - 1);
            node = block_node(P.n_grid, blk, local);
            const int plane = (int)(node / ((long long)P.n_grid * P.n_grid));
            owned = plane >= own_lo && plane < own_hi;
            if (halo.on == 1 || (halo.on == 3 && pass == 1)) {
                Vec4<T> q;
                if (halo_fetch<T>(halo, P.n_grid, blk, local, hs, q)) {
                    const Vec4<T> v = g_out[node];
                    g_out[node] = mk4<T>(v.x + q.x, v.y + q.y, v.z + q.z, v.w + q.w);
                }
            }
        }
        t_grid_bwd_node<T>(act, owned, node, threadIdx.x & 31, P, prims, s0, s1, grid_in, g_out, g_in, clear != 0, prim_grad, pf);
    }
    }
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_grid_fwd(SimConst<T> P, PrimSet<T> prims, const double* traj, SlotRef pf,
                                                     Vec4<T>* grid_in, Vec4<T>* grid_out, int clear_in, long long n_nodes) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    load_poses_smem<T>(traj, pf.get(), P.n_prim, s0, s1);
    long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_nodes) return;
    grid_fwd_body<T>(node, P, prims, s0, s1, grid_in, grid_out, clear_in != 0);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_g2p(SimConst<T> P, T* frames, long long n_pad, SlotRef slot_in, SlotRef slot_out,
                                                const Vec4<T>* grid_out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    g2p_body<T>(p, P, frame_at(frames, slot_in.get(), n_pad), frame_at(frames, slot_out.get(), n_pad), grid_out);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_g2p_bwd(SimConst<T> P, T* frames, long long n_pad, SlotRef slot_in, T* adj_next,
                                                    T* adj_cur, const Vec4<T>* grid_out, Vec4<T>* g_out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    g2p_bwd_body<T>(p, P, frame_at(frames, slot_in.get(), n_pad), frame_at(adj_next, 0, n_pad), frame_at(adj_cur, 0, n_pad),
                    grid_out, g_out);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_grid_bwd(SimConst<T> P, PrimSet<T> prims, const double* traj, SlotRef pfr,
                                                     Vec4<T>* grid_in, Vec4<T>* g_out, Vec4<T>* g_in, int clear,
                                                     double* prim_grad, long long n_nodes) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    const int pf = pfr.get();
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    PoseGrad<T> g0[PLB_MAX_PRIM], g1[PLB_MAX_PRIM];
    unsigned touched = 0;
    if (node < n_nodes) {
        for (int k = 0; k < P.n_prim; k++) { g0[k].clear(); g1[k].clear(); }
        grid_bwd_body<T>(node, P, prims, s0, s1, grid_in, g_out, g_in, clear != 0, g0, g1, touched);
    }
    for (int k = 0; k < P.n_prim; k++) {
        unsigned any = __ballot_sync(0xffffffffu, (touched >> k) & 1u);
        if (any == 0) continue;
        if (!((touched >> k) & 1u)) { g0[k].clear(); g1[k].clear(); }
        reduce_pose_grad<T>(g0[k], prim_grad + ((long long)pf * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
        reduce_pose_grad<T>(g1[k], prim_grad + ((long long)(pf + 1) * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
    }
}

template <class T, bool kSvd>
__global__ void __launch_bounds__(kBlock, Occ<T>::p2g_bwd) k_p2g_bwd(SimConst<T> P, T* frames, long long n_pad, SlotRef slot_in, T* adj_next,
                                                    T* adj_cur, Material<T> mat, const Vec4<T>* g_in, T* svd_base) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    SvdPtr<T> sp = svd_at(svd_base, slot_in.get(), n_pad);
    p2g_bwd_body<T, kSvd>(p, P, frame_at(frames, slot_in.get(), n_pad), frame_at(adj_next, 0, n_pad), frame_at(adj_cur, 0, n_pad), mat, g_in, &sp);
}

// ------------------------------------------------------------------------------------------------ loss
// accumulator layout (doubles): [0] density  [1] sdf  [2] sum m*t  [3] sum m  [4] max m (bits)  [8+k] min_dist_k (bits)
constexpr int kAccN = 8 + 2 * PLB_MAX_PRIM;    // [8+k] hard: min distance | soft: sum d*sw (-> soft distance), [16+k] soft: sum sw

template <class T>
__global__ void __launch_bounds__(kBlock) k_loss_mass(SimConst<T> P, T* frames, long long n_pad, int slot, T* grid_mass) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    loss_mass_body<T>(p, P, frame_at(frames, slot, n_pad), grid_mass);
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_loss_mass_tile(SimConst<T> P, T* frames, long long n_pad, int slot, T* grid_mass) {
    __shared__ T tiles[(kBlock / 32) * kTileVec4];
    t_loss_mass<T>(blockIdx.x * blockDim.x + threadIdx.x, threadIdx.x & 31, tiles + (threadIdx.x >> 5) * kTileVec4, P,
                   frame_at(frames, slot, n_pad), grid_mass);
}

__global__ void k_loss_init(double* acc, int soft) {
    int i = threadIdx.x;
    if (i < kAccN) acc[i] = (i >= 8 && i < 8 + PLB_MAX_PRIM && !soft) ? 100000.0 : 0.0;
}

template <class T>
__global__ void __launch_bounds__(256) k_loss_reduce(const T* __restrict__ grid_mass, const T* __restrict__ target,
                                                     const T* __restrict__ target_sdf, long long n_nodes, double* acc) {
    double d = 0, s = 0, mt = 0, sm = 0, mx = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += (long long)gridDim.x * blockDim.x) {
        double m = (double)grid_mass[i], t = (double)target[i];
        d += fabs(m - t);
        s += (double)target_sdf[i] * m;
        mt += m * t;
        sm += m;
        mx = fmax(mx, m);
    }
    d = warp_sum(d); s = warp_sum(s); mt = warp_sum(mt); sm = warp_sum(sm);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) {
        if (d != 0.0) atomicAdd(acc + 0, d);
        if (s != 0.0) atomicAdd(acc + 1, s);
        if (mt != 0.0) atomicAdd(acc + 2, mt);
        if (sm != 0.0) atomicAdd(acc + 3, sm);
        atomicMax(reinterpret_cast<unsigned long long*>(acc + 4), (unsigned long long)__double_as_longlong(mx));
    }
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_loss_contact(SimConst<T> P, PrimSet<T> prims, const double* traj, int pf,
                                                         T* frames, long long n_pad, int slot, double* acc, int soft) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = p < P.n_particles;
    V3<T> x = ok ? load_x(frame_at(frames, slot, n_pad), p) : zero3<T>();
    for (int k = 0; k < P.n_prim; k++) {
        if (!prims.s[k].movable) continue;
        if (soft) {
            // soft minimum (loss.py:112-135): sum_i sw(d_i) and sum_i d_i sw(d_i), sw(d) = 1 / (1 + 1e4 d^2)
            double d = ok ? (double)tmax(prim_sdf(prims.s[k], s0[k], x), T(0)) : 0.0;
            double sw = ok ? 1.0 / (1.0 + d * d * 10000.0) : 0.0;
            double a = warp_sum(sw), b = warp_sum(d * sw);
            if ((threadIdx.x & 31) == 0) { atomicAdd(acc + 8 + PLB_MAX_PRIM + k, a); if (b != 0.0) atomicAdd(acc + 8 + k, b); }
        } else {
            double d = 1e30;
            if (ok) d = (double)tmax(tmax(prim_sdf(prims.s[k], s0[k], x), T(0)), T(0));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d = fmin(d, __shfl_xor_sync(0xffffffffu, d, o));
            if ((threadIdx.x & 31) == 0)
                atomicMin(reinterpret_cast<unsigned long long*>(acc + 8 + k), (unsigned long long)__double_as_longlong(d));
        }
    }
}

// One thread: folds the accumulators into the step loss and the running loss; writes a record of 8 doubles.
template <class T>
__global__ void k_loss_finalize(PrimSet<T> prims, int n_prim, LossWeights w, const double* acc, double target_max,
                                double target_sum, double* loss_total, double* record) {
    double contact = 0;
    for (int k = 0; k < n_prim; k++)
        if (prims.s[k].movable) {
            double md = w.soft ? acc[8 + k] / acc[8 + PLB_MAX_PRIM + k] : acc[8 + k];
            contact += md * md;
        }
    double step = contact * w.contact + acc[0] * w.density + acc[1] * w.sdf;
    *loss_total += step;
    double ma = __longlong_as_double((long long)reinterpret_cast<const unsigned long long*>(acc)[4]);
    double I = acc[2] / ma / target_max;
    double U = acc[3] / ma + target_sum / target_max;
    record[0] = *loss_total; record[1] = contact; record[2] = acc[0]; record[3] = acc[1];
    record[4] = I / (U - I); record[5] = step; record[6] = 0; record[7] = 0;
}

template <class T>
__global__ void __launch_bounds__(kBlock) k_loss_bwd(SimConst<T> P, PrimSet<T> prims, const double* traj, int pf, T* frames,
                                                     long long n_pad, int slot, T* adj, const T* grid_mass, const T* target,
                                                     const T* target_sdf, LossWeights w, const double* acc, int contact_all,
                                                     double* prim_grad) {
    __shared__ Pose<T> s0[PLB_MAX_PRIM], s1[PLB_MAX_PRIM];
    load_poses_smem<T>(traj, pf, P.n_prim, s0, s1);
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    PoseGrad<T> g[PLB_MAX_PRIM];
    unsigned touched = 0;
    if (p < P.n_particles) {
        for (int k = 0; k < P.n_prim; k++) g[k].clear();
        loss_bwd_body<T>(p, P, frame_at(frames, slot, n_pad), frame_at(adj, 0, n_pad), grid_mass, target, target_sdf,
                         (T)w.sdf, (T)w.density, (T)w.contact, prims, s0, acc + 8, contact_all, g, touched, w.soft, acc + 8 + PLB_MAX_PRIM);
    }
    for (int k = 0; k < P.n_prim; k++) {
        unsigned any = __ballot_sync(0xffffffffu, (touched >> k) & 1u);
        if (any == 0) continue;
        if (!((touched >> k) & 1u)) g[k].clear();
        reduce_pose_grad<T>(g[k], prim_grad + ((long long)pf * PLB_MAX_PRIM + k) * PLB_POSE_DIM);
    }
}

// ------------------------------------------------------------------------------------------------ target SDF
// One Jacobi sweep of Loss.update_target_sdf (plb/engine/losses/loss.py:81-101): every node looks at the 6^3-1
// neighbours with offsets in [-3,3)^3 in lexicographic order and keeps the nearest propagated surface point.
__global__ void __launch_bounds__(128) k_sdf_sweep(int n, double dx, const double* __restrict__ density,
                                                   const double* __restrict__ sdf_c, const double* __restrict__ near_c,
                                                   double* sdf_o, double* near_o, int* changed) {
    long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n * n * n;
    if (node >= total) return;
    const unsigned un = (unsigned)node, nn = (unsigned)n, row = un / nn;        // n_grid <= 1024: node < 2^30, 32-bit divisions
    int k = (int)(un - row * nn), i = (int)(row / nn), j = (int)(row - (unsigned)i * nn);
    const double inf = 1000.0;
    double gx = i * dx, gy = j * dx, gz = k * dx;
    double best = inf;
    double bx = near_c[node * 3 + 0], by = near_c[node * 3 + 1], bz = near_c[node * 3 + 2];
    if (density[node] > 1e-4) {
        best = 0.0; bx = gx; by = gy; bz = gz;
    } else {
        for (int a = -3; a < 3; a++)
            for (int b = -3; b < 3; b++)
                for (int c = -3; c < 3; c++) {
                    int vi = i + a, vj = j + b, vk = k + c;
                    if (vi < 0 || vj < 0 || vk < 0 || vi >= n || vj >= n || vk >= n) continue;
                    if (a == 0 && b == 0 && c == 0) continue;
                    long long v = ((long long)vi * n + vj) * n + vk;
                    if (sdf_c[v] < inf) {
                        double px = near_c[v * 3 + 0], py = near_c[v * 3 + 1], pz = near_c[v * 3 + 2];
                        double ddx = gx - px, ddy = gy - py, ddz = gz - pz;
                        double dist = sqrt(ddx * ddx + ddy * ddy + ddz * ddz + 1e-8);
                        if (dist < best) { best = dist; bx = px; by = py; bz = pz; }
                    }
                }
    }
    if (best != sdf_c[node] || bx != near_c[node * 3] || by != near_c[node * 3 + 1] || bz != near_c[node * 3 + 2]) *changed = 1;
    sdf_o[node] = best;
    near_o[node * 3 + 0] = bx; near_o[node * 3 + 1] = by; near_o[node * 3 + 2] = bz;
}

// ------------------------------------------------------------------------------------------------ host <-> frame
// AoS float64 host layout (x[N][3], v[N][3], F[N][3][3], C[N][3][3], staged on the device) <-> packed planes
// `perm[p]` = caller-side (reference order) index of the particle stored at position p (identity until the first sort)
template <class T>
__global__ void k_pack_frame(int n, long long n_pad, T* frame, const int* __restrict__ perm, const double* x, const double* v,
                             const double* F, const double* C) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long q = perm[p];
    FramePtr<T> f = frame_at(frame, 0, n_pad);
    V3<T> xx, vv; M3<T> CC;
    load_xvC(f, p, xx, vv, CC);
    M3<T> FF = load_F(f, p);
    if (x) xx = mk3<T>((T)x[q * 3], (T)x[q * 3 + 1], (T)x[q * 3 + 2]);
    if (v) vv = mk3<T>((T)v[q * 3], (T)v[q * 3 + 1], (T)v[q * 3 + 2]);
    if (C) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) CC.m[i][j] = (T)C[q * 9 + i * 3 + j];
    if (F) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) FF.m[i][j] = (T)F[q * 9 + i * 3 + j];
    store_xvC(f, p, xx, vv, CC);
    store_F(f, p, FF);
}

template <class T>
__global__ void k_unpack_frame(int n, long long n_pad, T* frame, const int* __restrict__ perm, double* x, double* v, double* F,
                               double* C) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long q = perm[p];
    FramePtr<T> f = frame_at(frame, 0, n_pad);
    V3<T> xx, vv; M3<T> CC;
    load_xvC(f, p, xx, vv, CC);
    M3<T> FF = load_F(f, p);
    if (x) { x[q * 3] = xx.x; x[q * 3 + 1] = xx.y; x[q * 3 + 2] = xx.z; }
    if (v) { v[q * 3] = vv.x; v[q * 3 + 1] = vv.y; v[q * 3 + 2] = vv.z; }
    if (C) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[q * 9 + i * 3 + j] = CC.m[i][j];
    if (F) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F[q * 9 + i * 3 + j] = FF.m[i][j];
}

// ---- spatial sort support: key = (4^3-block id, cell within block) of the particle's base cell
template <class T>
__global__ void k_sort_keys(SimConst<T> P, T* frame, long long n_pad, unsigned* keys, int* vals) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_particles) return;
    V3<T> x = load_x(frame_at(frame, 0, n_pad), p);
    int b[3];
#pragma unroll
    for (int d = 0; d < 3; d++) b[d] = (int)(x[d] * P.inv_dx - T(0.5));
    const int nbx = P.n_grid >> kBlkShift;
    unsigned blk = (unsigned)block_id(nbx, b[0], b[1], b[2]);
    unsigned cell = (unsigned)(((b[0] & 3) << 4) | ((b[1] & 3) << 2) | (b[2] & 3));
    keys[p] = blk * 64u + cell;
    vals[p] = p;
}
// dst[j] = src[order[j]] for a whole frame; perm_out[j] = perm_in[order[j]]
template <class T>
__global__ void k_permute_frame(int n, long long n_pad, const T* src, T* dst, const int* __restrict__ order,
                                const int* __restrict__ perm_in, int* perm_out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int o = order[j];
    FramePtr<T> fs = frame_at(const_cast<T*>(src), 0, n_pad), fd = frame_at(dst, 0, n_pad);
    V3<T> x, v; M3<T> C;
    load_xvC(fs, o, x, v, C);
    M3<T> F = load_F(fs, o);
    store_xvC(fd, j, x, v, C);
    store_F(fd, j, F);
    if (perm_out) perm_out[j] = perm_in[o];
}
template <class T> __global__ void k_permute_scalar(int n, const T* src, T* dst, const int* __restrict__ order) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[order[j]];
}
// the inverse: dst[order[j]] = src[j] (env-step re-sort: the adjoint frame and the boundary frame go back to the previous env step's order)
template <class T>
__global__ void k_unpermute_frame(int n, long long n_pad, const T* src, T* dst, const int* __restrict__ order) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int o = order[j];
    FramePtr<T> fs = frame_at(const_cast<T*>(src), 0, n_pad), fd = frame_at(dst, 0, n_pad);
    V3<T> x, v; M3<T> C;
    load_xvC(fs, j, x, v, C);
    M3<T> F = load_F(fs, j);
    store_xvC(fd, o, x, v, C);
    store_F(fd, o, F);
}
template <class T> __global__ void k_unpermute_scalar(int n, const T* src, T* dst, const int* __restrict__ order) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[order[j]] = src[j];
}
// ---- policy path: observation gather / observation-adjoint scatter for a handful of particles (caller-order indices)
__global__ void k_invert_perm(int n, const int* __restrict__ perm, int* inv) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) inv[perm[p]] = p;
}
template <class T>
__global__ void k_gather_xv(int n_sel, long long n_pad, T* frame, const int* __restrict__ idx, const int* __restrict__ inv, double* out6) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_sel) return;
    const int p = inv[idx[j]];
    FramePtr<T> f = frame_at(frame, 0, n_pad);
    Vec4<T> q0 = f.A0[p], q1 = f.A1[p];
    out6[j * 6 + 0] = q0.x; out6[j * 6 + 1] = q0.y; out6[j * 6 + 2] = q0.z;
    out6[j * 6 + 3] = q0.w; out6[j * 6 + 4] = q1.x; out6[j * 6 + 5] = q1.y;
}
template <class T>
__global__ void k_scatter_adj_xv(int n_sel, long long n_pad, T* adj, const int* __restrict__ idx, const int* __restrict__ inv, const double* g6) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_sel) return;
    const int p = inv[idx[j]];                  // distinct indices: no two threads touch the same particle
    FramePtr<T> f = frame_at(adj, 0, n_pad);
    Vec4<T> q0 = f.A0[p], q1 = f.A1[p];
    q0.x += (T)g6[j * 6 + 0]; q0.y += (T)g6[j * 6 + 1]; q0.z += (T)g6[j * 6 + 2];
    q0.w += (T)g6[j * 6 + 3]; q1.x += (T)g6[j * 6 + 4]; q1.y += (T)g6[j * 6 + 5];
    f.A0[p] = q0; f.A1[p] = q1;
}
__global__ void k_iota(int n, int* a) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }

template <class T> __global__ void k_convert(long long n, const double* src, T* dst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (T)src[i];
}
template <class T> __global__ void k_convert_perm(int n, const double* src, T* dst, const int* __restrict__ perm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (T)src[perm[i]];
}
template <class T> __global__ void k_grid_to_double(long long n, const Vec4<T>* src, double* dst) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { Vec4<T> v = src[i]; dst[i * 4] = v.x; dst[i * 4 + 1] = v.y; dst[i * 4 + 2] = v.z; dst[i * 4 + 3] = v.w; }
}
template <class T> __global__ void k_count_active(long long n, const Vec4<T>* grid_in, unsigned long long* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool a = (i < n) && (grid_in[i].w > T(1e-12));
    unsigned m = __ballot_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

}  // namespace plb

#include "plb_tile.cuh"
