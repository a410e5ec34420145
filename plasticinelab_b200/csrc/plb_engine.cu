// Engine object + C ABI (include/plb_b200.h).  Owns every device buffer, the primitive trajectories (host f64,
// mirrored on the device) and the launch sequence of a forward / backward substep.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <map>
#include <tuple>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include "../../include/plb_b200.h"
#include "plb_kernels.cuh"
#include "plb_kinematics.hpp"
#include "plb_setup.hpp"

using namespace plb;

static std::string g_create_error;

#define PLB_CUDA(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            err = std::string(#expr) + ": " + cudaGetErrorString(_e);                                \
            return PLB_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

#define PLB_REQUIRE(cond, msg)                                                                       \
    do {                                                                                             \
        if (!(cond)) { err = std::string(msg) + " [" #cond "]"; return PLB_ERR_INVALID; }            \
    } while (0)

enum KernelId { K_P2G = 0, K_GRID_FWD, K_G2P, K_P2G_RECOMPUTE, K_GRID_FWD_RECOMPUTE, K_G2P_BWD, K_GRID_BWD, K_P2G_BWD,
                K_LOSS_FWD, K_LOSS_BWD, K_MISC, K_G2P_P2G, K_P2G_BWD_G2P_BWD, K_COUNT };
static const char* kKernelNames[K_COUNT] = {"p2g", "grid_fwd", "g2p", "p2g_recompute", "grid_fwd_recompute", "g2p_bwd", "grid_bwd",
                                            "p2g_bwd", "loss_fwd", "loss_bwd", "misc", "g2p_p2g", "p2g_bwd_g2p_bwd"};

struct plb_engine {
    std::string err;
    // optional per-launch CUDA-event timing (bench.py's live roofline numbers); off by default
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_kid;
    size_t prof_used = 0;
    double prof_ms[K_COUNT] = {0};
    long long prof_cnt[K_COUNT] = {0};
    cudaStream_t prof_stream = 0;
    cudaStream_t prof_cur = 0;
    // Programmatic dependent launch of the kernels chained inside the env-step graphs (plb_types.cuh, pdl_wait / pdl_launch):
    // the launch may become resident while the kernel ahead of it in the stream drains.  Off while profiling (the event pairs
    // around single kernels should not overlap) and in slab runs (PLB_PDL=0 switches it off altogether).
    bool pdl_enable = true, pdl_inhibit = false;          // (inhibit: slab runs whose halo needs a launch without pdl_wait between the chained kernels)
    bool pdl_on() const { return pdl_enable && !prof_on && !pdl_inhibit; }
    template <class... KArgs, class... Args>
    void launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&lc, kern, KArgs(std::forward<Args>(args))...);
    }
    void prof_begin(int kid, cudaStream_t st = nullptr) {
        if (!prof_on) return;
        prof_cur = st ? st : prof_stream;
        if (prof_used * 2 + 2 > prof_ev.size()) {
            size_t old = prof_ev.size();
            prof_ev.resize(old + 8192);
            for (size_t i = old; i < prof_ev.size(); i++) cudaEventCreate(&prof_ev[i]);
        }
        if (prof_kid.size() <= prof_used) prof_kid.resize(prof_used + 4096);
        prof_kid[prof_used] = kid;
        cudaEventRecord(prof_ev[prof_used * 2], prof_cur);
    }
    void prof_end() {
        if (!prof_on) return;
        cudaEventRecord(prof_ev[prof_used * 2 + 1], prof_cur);
        prof_used++;
    }
    void prof_collect() {
        if (prof_used) cudaDeviceSynchronize();          // (launches may sit on the engine's side stream too)
        for (size_t i = 0; i < prof_used; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, prof_ev[i * 2], prof_ev[i * 2 + 1]);
            prof_ms[prof_kid[i]] += ms;
            prof_cnt[prof_kid[i]]++;
        }
        prof_used = 0;
    }
    virtual ~plb_engine() {}
    virtual int init(const plb_config& c, const plb_primitive_desc* prims) = 0;
    virtual int set_materials(const double* mu, const double* lam, const double* ys) = 0;
    virtual int set_frame(int slot, const double* x, const double* v, const double* F, const double* C) = 0;
    virtual int get_frame(int slot, double* x, double* v, double* F, double* C) = 0;
    virtual int copy_frame(int src, int dst) = 0;
    virtual int sort_particles(int slot) = 0;
    virtual int frame_ptr(int slot, void** ptr, long long* n_pad, int* sb) = 0;
    virtual int set_prim_state(int pf, int k, const double* s) = 0;
    virtual int get_prim_state(int pf, int k, double* s) = 0;
    virtual int copy_prim_frame(int src, int dst) = 0;
    virtual int set_softness(double s) = 0;
    virtual int set_action(int step, int S, const double* a, int n) = 0;
    virtual int kinematics(int pf, int n) = 0;
    virtual int substep_fwd(int si, int so, int pf) = 0;
    virtual int substep_bwd(int si, int pf) = 0;
    virtual int step_fwd(int slot0, int pf0, int n) = 0;
    virtual int step_bwd(int slot0, int pf0, int n) = 0;
    virtual int zero_grads() = 0;
    virtual int set_adjoint(const double* gx, const double* gv, const double* gF, const double* gC) = 0;
    virtual int get_adjoint(double* gx, double* gv, double* gF, double* gC) = 0;
    virtual int get_prim_grads(int pf0, int n, double* out) = 0;
    virtual int get_action_grad(int n_steps, int S, double* out) = 0;
    virtual int action_grad_step(int step, int S, double* out) = 0;
    virtual int add_pose_adjoint(int k, const double* g8) = 0;
    virtual int gather_particles(int slot, const int* idx, int n, double* x3, double* v3) = 0;
    virtual int scatter_adjoint(const int* idx, int n, const double* gx3, const double* gv3) = 0;
    virtual int set_target(const double* density, const double* sdf) = 0;
    virtual int get_target_sdf(double* sdf) = 0;
    virtual int set_loss_weights(double sdf, double density, double contact, int soft, int all) = 0;
    virtual int loss_fwd(int slot, int pf, double* out8) = 0;
    virtual int loss_bwd(int slot, int pf) = 0;
    virtual int get_loss(double* v) = 0;
    virtual int clear_loss() = 0;
    virtual int debug_get_grid(double* in4, double* out4) = 0;
    virtual int count_active(int slot, long long* n) = 0;
    virtual int slab_configure(int own_lo, int own_hi, int halo_w, int has_left, int has_right) = 0;
    virtual int slab_buffer(int which, int side, int dir, void** ptr, long long* bytes) = 0;
    virtual int slab_fwd_p2g(int si, int so) = 0;
    virtual int slab_fwd_finish(int si, int so, int pf) = 0;
    virtual int slab_bwd_begin(int si, int pf) = 0;
    virtual int slab_bwd_finish(int si, int pf) = 0;
    virtual int slab_loss_begin(int slot) = 0;
    virtual int slab_loss_reduce(int slot, int pf) = 0;
    virtual int slab_loss_finish(int slot, int pf, int backward, double* out8) = 0;
    virtual int device_buffer(int which, void** ptr, long long* bytes) = 0;
    virtual int slab_ipc_export(int side, void* handle64) = 0;
    virtual int slab_ipc_import(int side, const void* handle64) = 0;
    virtual int slab_ipc_close() = 0;
    virtual int slab_ipc_export_grid(int which, void* handle64) = 0;
    virtual int slab_ipc_import_grid(int side, int which, const void* handle64) = 0;
    virtual int set_stream(void* s) = 0;
    virtual int synchronize() = 0;
    long long launches = 0;
};

// __launch_bounds__ min-blocks instantiated for the fused particle kernels (float: two register caps each; double: one)
template <class T> struct OccSel { static constexpr int fwd_lo = 1, fwd_hi = 1, bwd_lo = 1, bwd_hi = 1; };
template <> struct OccSel<float> { static constexpr int fwd_lo = 5, fwd_hi = 6, bwd_lo = 3, bwd_hi = 4; };

// same for the chunked kernels (plb_tile.cuh): 28.5 KB of shared memory per CTA, so the register cap decides the resident CTAs
template <class T> struct TileOcc { static constexpr int fwd = 1, fwd_hi = 1, bwd_lo = 1, bwd_hi = 1; };
template <> struct TileOcc<float> { static constexpr int fwd = 5, fwd_hi = 6, bwd_lo = 3, bwd_hi = 4; };

template <class T>
struct Engine : plb_engine {
    plb_config cfg{};
    SimConst<T> P{};
    PrimSet<T> prims{};
    std::vector<kin::Desc> kdesc;
    std::vector<plb_primitive_desc> pdesc;
    cudaStream_t stream = 0;
    long long n_pad = 0, n_nodes = 0;
    int action_total = 0;
    std::vector<int> action_off;

    // device buffers
    T* frames = nullptr;            // [max_frames][24][n_pad]
    T* adj[2] = {nullptr, nullptr}; // ping-pong adjoint frames; adj[cur] holds the adjoint of frame `adj_frame`
    int cur = 0;
    Vec4<T>* grid_in = nullptr;     // (momentum, mass); zero between substeps
    Vec4<T>* grid_out = nullptr;    // velocity after the grid operator
    Vec4<T>* g_out = nullptr;       // adjoint of grid_out; zero between substeps
    Vec4<T>* g_in = nullptr;        // adjoint of grid_in
    T* mat_mu = nullptr; T* mat_lam = nullptr; T* mat_ys = nullptr;
    T* grid_mass = nullptr; T* target = nullptr; T* target_sdf = nullptr;
    double* d_stage = nullptr;      // staging for host<->device f64 AoS frames: 24 * n doubles
    double* d_traj = nullptr;       // [max_prim_frames][PLB_MAX_PRIM][8]
    double* d_prim_grad = nullptr;  // same shape
    double* d_acc = nullptr;        // loss accumulators (kAccN) + [kAccN] running loss + [kAccN+1 ..] record(8)
    unsigned long long* d_count = nullptr;
    // active 4^3 blocks of the current substep
    unsigned char* d_flags = nullptr; int* d_list = nullptr; int* d_nactive = nullptr; int n_blocks = 0;
    bool sparse = true;
    // forward-grid store + CUDA graphs
    GridStore<T> store{nullptr, nullptr, nullptr, nullptr, 0};
    std::vector<char> stored;       // host view: slot s holds the grid of the substep that started at s
    int* d_cursor = nullptr;
    bool use_graphs = true;
    cudaStream_t own_stream = nullptr;
    struct GraphKey { int dir, n, parity, stored; bool operator<(const GraphKey& o) const {
        return std::tie(dir, n, parity, stored) < std::tie(o.dir, o.n, o.parity, o.stored); } };
    std::map<GraphKey, cudaGraphExec_t> graphs;
    std::map<GraphKey, long long> graph_nodes;
    // slab decomposition (multi-GPU): owned planes [own_lo, own_hi), zones of +-halo_w planes around the boundaries
    struct Slab { bool on = false; int own_lo = 0, own_hi = 0, w = 0; bool has[2] = {false, false}; int zlo[2] = {0, 0}, zhi[2] = {0, 0};
                  void* recv[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
                  // peer-memory halo: my inboxes (neighbours write here), the neighbours' inboxes mapped through CUDA IPC
                  char* inbox[2] = {nullptr, nullptr}; char* peer[2] = {nullptr, nullptr}; HaloGeom geom[2]; size_t inbox_bytes = 0;
                  int* seq = nullptr; int* listed_stamp = nullptr; int* err = nullptr; unsigned* done = nullptr; bool peer_ready = false; bool exported = false, ipc_closed = false;
                  bool fused = true;
                  // direct halo (PLB_SLAB_DIRECT=0 disables): scatter kernels RED their zone contributions into the neighbours' grids
                  // (CUDA-IPC mappings of grid_in x2 and g_out x2, alternating by substep parity); an exchange is a completion flag
                  bool direct = false; Vec4<T>* g_out2 = nullptr; Vec4<T>* peer_grid[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}}; } slab;      // fused: env-step block list + one push launch per exchange, receive inside the grid kernels (PLB_SLAB_FUSED=0: the per-substep chain)
    bool tile_scatter = true;       // warp-tile pre-reduced scatters (kernel_variant 0); variant 2 = sparse grid + direct atomics
    bool fuse = true;               // fused G2P+P2G / P2G.grad+G2P.grad particle kernels inside env-step graphs (PLB_FUSE=0 disables)
    static constexpr int cta = kBlock;          // threads per CTA of the per-warp scatter kernels
    int fwd_minb = 5, bwd_minb = 4; // register caps of the fused particle kernels (OccSel; PLB_FWD_MINB=6 selects the tighter forward cap, PLB_BWD_MINB=3 the looser backward one)
    int flush_mode = 3;             // full-tile flush of the per-warp kernels: 3 = runs of consecutive lanes (flush_runs, default), 0 = per-cell groups (PLB_FLUSH_MODE=0)
    bool grid_bwd_v2 = true;        // grid adjoint with register-resident pose gradients (k_grid_bwd_sparse_v2); PLB_GRID_BWD_V2=0: array form
    bool env_list = true;           // active-block list built once per env step (dilated by one block) instead of per substep (PLB_ENV_LIST=0: per substep)
    unsigned char* d_flags2 = nullptr; unsigned char* d_listed = nullptr;
    bool bwd_overlap = true;        // backward graphs: restore + grid recompute of substep s-1 on a forked branch, overlapping the
                                    // particle kernel and grid adjoint of substep s (needs the second grid set; PLB_BWD_OVERLAP=0 disables)
    // grid set: forward grid (momentum+mass), grid operator output, active-block list.  Set 0 is the working set of the
    // forward pass; the backward pass alternates between the two sets by substep parity when bwd_overlap is on.
    struct GridSet { Vec4<T>* in; Vec4<T>* out; int* list; int* count; };
    GridSet sets[2] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    cudaStream_t side_stream = nullptr;
    std::vector<cudaEvent_t> cap_events; size_t cap_ev_used = 0;
    cudaEvent_t next_event() {
        if (cap_ev_used == cap_events.size()) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); cap_events.push_back(e); }
        return cap_events[cap_ev_used++];
    }
    std::vector<char> fwd_ok;       // host view: slot s+1 holds the frame the forward substep produced from slot s
    // SVD store (PLB_SVD_STORE=1): U, sigma, V of F_tmp per particle and frame slot, written by the forward P2G, read by the
    // backward pass instead of re-running the Jacobi iteration; 84 B (f32) per particle and frame
    T* svd_store = nullptr; bool svd_enable = true;
    std::vector<char> svd_ok;       // host view: svd_store[slot] holds the decomposition of the substep that started at slot
    // spatial sort: d_perm[p] = caller-side index of the particle stored at position p
    int* d_perm = nullptr; int* d_perm2 = nullptr; unsigned* d_keys = nullptr; unsigned* d_keys2 = nullptr;
    int* d_vals = nullptr; int* d_vals2 = nullptr; void* d_cub = nullptr; size_t cub_bytes = 0; T* frame_tmp = nullptr;
    // chunked particle kernels with TMA-loaded grid windows (plb_tile.cuh) inside the env-step graphs; PLB_TILE=0: the
    // per-thread-gather kernels.  The chunk table indexes the sorted order: rebuilt by every plb_sort_particles.
    bool tile_mode = true;
    // backward chunk kernels (PLB_TILE_BWD=1): measured slower than the per-thread-gather backward kernel on a B200 (171 vs 159 us
    // at 1M particles: the chunk -> TMA -> wait chain at CTA start and instruction-fetch stalls outweigh the shared-memory
    // gathers in a kernel that registers cap at 16 warps per SM either way), so the default backward path keeps the latter
    // Env-step re-sort (default; PLB_RESORT=0 switches it off): plb_step_fwd re-sorts the particles of its first frame by (block, cell)
    // when that frame was produced by an earlier env step, so that a moving body keeps few distinct cells per warp (the scatter cost:
    // +41 % on a translating body, +4 % on a 10-step Move episode at 1M particles, neutral at 100k; profiles/r2h_resort_timeline.txt).
    // The permutation q of every re-sorted frame is kept; plb_step_bwd, once it has produced the adjoint of that frame, puts the
    // adjoint frame, the frame itself and the materials back into the previous env step's order.  order_perm[k] = caller-side index
    // map of ordering k (0 = the ordering of the last plb_sort_particles), slot_order[s] = ordering frame s is stored in.
    // Not used in slab runs, with the chunked backward kernels, or on frames that were not produced by plb_step_fwd.
    bool resort = true;
    std::map<int, int*> resort_q;
    std::vector<int*> q_spent;          // free list of permutation buffers (stream-ordered reuse)
    int orders_valid = 0;               // order_perm[0 .. orders_valid) hold orderings of the current episode
    std::vector<int*> order_perm;
    std::vector<int> slot_order;
    int cur_order = 0;
    bool window_follow = true;      // chunk window origins recomputed from the first frame of every env step (PLB_WINDOW_FOLLOW=0: fixed at the sort)
    bool tile_bwd = false;
    int tile_fwd_minb = 6;          // chunked forward kernel: 6 resident CTAs per SM (80 registers, no spills) | 5 (96 registers): PLB_TILE_FWD_MINB
    bool svd_warm = true;           // Jacobi SVD of substep s+1 started from V of substep s (needs the SVD store; PLB_SVD_WARM=0: from the identity)
    Chunk* d_chunks = nullptr; int* d_nchunks = nullptr; int chunk_cap = 0, chunk_grid = 0;
    CUtensorMap tm_out[2], tm_gin;  // grid_out of the two grid sets, g_in
    ChunkTable chunk_table() const { ChunkTable t; t.chunks = d_chunks; t.n_chunks = d_nchunks; return t; }
    bool has_target = false;
    double target_max = 0, target_sum = 0;
    LossWeights lw{10.0, 10.0, 1.0, 0};
    int contact_all = 1;

    // host mirrors
    std::vector<double> traj;       // poses
    std::vector<double> vel;        // per frame per prim: v(3) w(3) gv(1) pad -> 8
    std::vector<double> actions;    // per env step: action_total
    int max_steps_actions = 0;

    ~Engine() override {
        cudaFree(frames); cudaFree(adj[0]); cudaFree(adj[1]); cudaFree(grid_in); cudaFree(grid_out); cudaFree(g_out);
        cudaFree(g_in); cudaFree(mat_mu); cudaFree(mat_lam); cudaFree(mat_ys); cudaFree(grid_mass); cudaFree(target);
        cudaFree(target_sdf); cudaFree(d_stage); cudaFree(d_traj); cudaFree(d_prim_grad); cudaFree(d_acc); cudaFree(d_count);
        cudaFree(d_flags); cudaFree(d_list); cudaFree(d_nactive); cudaFree(d_perm); cudaFree(d_perm2); cudaFree(d_keys);
        cudaFree(d_keys2); cudaFree(d_vals); cudaFree(d_vals2); cudaFree(d_cub); cudaFree(frame_tmp);
        cudaFree(d_chunks); cudaFree(d_nchunks);
        for (auto& kv : resort_q) cudaFree(kv.second);
        for (int* p : q_spent) cudaFree(p);
        for (int* p : order_perm) cudaFree(p);
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        cudaFree(store.vals); cudaFree(store.ids); cudaFree(store.cnt); cudaFree(store.overflow); cudaFree(d_cursor);
        cudaFree(d_flags2); cudaFree(d_listed);
        cudaFree(d_inv_perm); cudaFree(d_sel_idx); cudaFree(d_sel_val); cudaFree(svd_store);
        cudaFree(sets[1].in); cudaFree(sets[1].out); cudaFree(sets[1].list); cudaFree(sets[1].count);
        for (int side = 0; side < 2; side++) {
            // an exported inbox may only be freed once every importer has unmapped it: after plb_slab_ipc_close + a barrier
            // (engine/sharded.py close()); otherwise it is left to process exit
            if (!slab.peer[side] && (slab.ipc_closed || !slab.exported)) cudaFree(slab.inbox[side]);
            for (int which = 0; which < 3; which++) cudaFree(slab.recv[which][side]);
        }
        cudaFree(slab.seq); cudaFree(slab.err); cudaFree(slab.listed_stamp); cudaFree(slab.done);
        if (slab.ipc_closed || !slab.exported) cudaFree(slab.g_out2);
        for (cudaEvent_t e : cap_events) cudaEventDestroy(e);
        if (side_stream) cudaStreamDestroy(side_stream);
        if (own_stream) cudaStreamDestroy(own_stream);
    }

    int blocks(long long n, int b = kBlock) const { return (int)((n + b - 1) / b); }
    static size_t tile_smem_bytes(int threads) { return (size_t)(threads / 32) * kTileVec4 * sizeof(Vec4<T>); }
    Material<T> material() const { Material<T> m; m.mu = mat_mu; m.lam = mat_lam; m.ys = mat_ys; return m; }
    T* frame_base(int slot) const { return frames + (long long)slot * 24 * n_pad; }

    int init(const plb_config& c, const plb_primitive_desc* pd) override {
        cfg = c;
        PLB_REQUIRE(c.n_particles > 0 && c.n_grid >= 8 && c.max_frames >= 2 && c.max_prim_frames >= 2, "bad sizes");
        PLB_REQUIRE(c.n_primitives >= 0 && c.n_primitives <= PLB_MAX_PRIM, "too many primitives");
        PLB_CUDA(cudaSetDevice(c.device));
        n_pad = ((long long)c.n_particles + 31) / 32 * 32;
        n_nodes = (long long)c.n_grid * c.n_grid * c.n_grid;
        P = make_simconst<T>(c);

        action_off.assign(1, 0);
        traj.assign((size_t)c.max_prim_frames * PLB_MAX_PRIM * 8, 0.0);
        vel.assign((size_t)c.max_prim_frames * PLB_MAX_PRIM * 8, 0.0);
        for (int k = 0; k < c.n_primitives; k++) {
            const plb_primitive_desc& d = pd[k];
            pdesc.push_back(d);
            prims.s[k] = make_primstatic<T>(d, 0.0);
            kdesc.push_back(make_kindesc(d));
            action_off.push_back(action_off.back() + d.action_dim);
            for (int i = 0; i < 8; i++) traj[(size_t)k * 8 + i] = d.init_state[i];
        }
        action_total = action_off.back();
        max_steps_actions = c.max_prim_frames;
        actions.assign((size_t)max_steps_actions * std::max(action_total, 1), 0.0);

        size_t fbytes = (size_t)c.max_frames * 24 * n_pad * sizeof(T);
        PLB_CUDA(cudaMalloc(&frames, fbytes));
        PLB_CUDA(cudaMemset(frames, 0, fbytes));
        for (int i = 0; i < 2; i++) {
            PLB_CUDA(cudaMalloc(&adj[i], (size_t)24 * n_pad * sizeof(T)));
            PLB_CUDA(cudaMemset(adj[i], 0, (size_t)24 * n_pad * sizeof(T)));
        }
        size_t gbytes = (size_t)n_nodes * sizeof(Vec4<T>);
        PLB_CUDA(cudaMalloc(&grid_in, gbytes));  PLB_CUDA(cudaMemset(grid_in, 0, gbytes));
        PLB_CUDA(cudaMalloc(&grid_out, gbytes)); PLB_CUDA(cudaMemset(grid_out, 0, gbytes));
        PLB_CUDA(cudaMalloc(&g_out, gbytes));    PLB_CUDA(cudaMemset(g_out, 0, gbytes));
        PLB_CUDA(cudaMalloc(&g_in, gbytes));     PLB_CUDA(cudaMemset(g_in, 0, gbytes));
        PLB_CUDA(cudaMalloc(&grid_mass, n_nodes * sizeof(T)));
        PLB_CUDA(cudaMalloc(&target, n_nodes * sizeof(T)));       PLB_CUDA(cudaMemset(target, 0, n_nodes * sizeof(T)));
        PLB_CUDA(cudaMalloc(&target_sdf, n_nodes * sizeof(T)));   PLB_CUDA(cudaMemset(target_sdf, 0, n_nodes * sizeof(T)));
        PLB_CUDA(cudaMalloc(&d_stage, (size_t)24 * c.n_particles * sizeof(double)));
        size_t tb = traj.size() * sizeof(double);
        PLB_CUDA(cudaMalloc(&d_traj, tb));
        PLB_CUDA(cudaMemcpy(d_traj, traj.data(), tb, cudaMemcpyHostToDevice));
        PLB_CUDA(cudaMalloc(&d_prim_grad, tb));
        PLB_CUDA(cudaMemset(d_prim_grad, 0, tb));
        PLB_CUDA(cudaMalloc(&d_acc, (kAccN + 1 + 8) * sizeof(double)));
        PLB_CUDA(cudaMemset(d_acc, 0, (kAccN + 1 + 8) * sizeof(double)));
        PLB_CUDA(cudaMalloc(&d_count, sizeof(unsigned long long)));
        PLB_REQUIRE(c.n_grid % 4 == 0, "n_grid must be a multiple of 4");
        PLB_REQUIRE(c.n_grid <= 1024, "n_grid above 1024 is not supported (cell keys pack 10 bits per axis)");
        sparse = c.kernel_variant != 1;
        tile_scatter = c.kernel_variant == 0;
        if (const char* v = getenv("PLB_FWD_MINB")) fwd_minb = atoi(v);
        if (const char* v = getenv("PLB_BWD_MINB")) bwd_minb = atoi(v);
        if (const char* v = getenv("PLB_BWD_OVERLAP")) bwd_overlap = atoi(v) != 0;
        if (const char* v = getenv("PLB_SVD_STORE")) svd_enable = atoi(v) != 0;
        if (const char* v = getenv("PLB_ENV_LIST")) env_list = atoi(v) != 0;
        if (const char* v = getenv("PLB_FLUSH_MODE")) flush_mode = atoi(v) == 0 ? 0 : 3;
        if (const char* v = getenv("PLB_GRID_BWD_V2")) grid_bwd_v2 = atoi(v) != 0;
        fuse = c.kernel_variant == 0 && !(getenv("PLB_FUSE") && atoi(getenv("PLB_FUSE")) == 0);
        if (tile_scatter) {
            const int full = (int)tile_smem_bytes(kBlock);
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_warp<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, full));
            PLB_CUDA(cudaFuncSetAttribute(k_g2p_p2g_warp<T, OccSel<T>::fwd_lo>, cudaFuncAttributeMaxDynamicSharedMemorySize, full));
            PLB_CUDA(cudaFuncSetAttribute(k_g2p_p2g_warp<T, OccSel<T>::fwd_hi>, cudaFuncAttributeMaxDynamicSharedMemorySize, full));
            PLB_CUDA(cudaFuncSetAttribute(k_g2p_bwd_warp<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, full));
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_lo, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, full));
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_hi, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, full));
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_lo, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, full));
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_hi, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, full));
            // the full tiles want the whole shared-memory carve-out (4 x 57 KB per SM); the driver's default choice left the
            // backward kernel at 3 CTAs per SM by shared memory (ncu launch__occupancy_limit_shared_mem)
            const int carve = cudaSharedmemCarveoutMaxShared;
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_warp<T>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            PLB_CUDA(cudaFuncSetAttribute(k_g2p_p2g_warp<T, OccSel<T>::fwd_lo>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            PLB_CUDA(cudaFuncSetAttribute(k_g2p_p2g_warp<T, OccSel<T>::fwd_hi>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            PLB_CUDA(cudaFuncSetAttribute(k_g2p_bwd_warp<T>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_lo, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_hi, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_lo, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            PLB_CUDA(cudaFuncSetAttribute(k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_hi, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        }
        n_blocks = (c.n_grid / 4) * (c.n_grid / 4) * (c.n_grid / 4);
        PLB_CUDA(cudaMalloc(&d_flags, n_blocks));
        PLB_CUDA(cudaMemset(d_flags, 0, n_blocks));
        PLB_CUDA(cudaMalloc(&d_flags2, n_blocks));
        PLB_CUDA(cudaMemset(d_flags2, 0, n_blocks));
        PLB_CUDA(cudaMalloc(&d_listed, n_blocks));
        PLB_CUDA(cudaMemset(d_listed, 0, n_blocks));
        PLB_CUDA(cudaMalloc(&d_list, n_blocks * sizeof(int)));
        PLB_CUDA(cudaMalloc(&d_nactive, sizeof(int)));
        PLB_CUDA(cudaMemset(d_nactive, 0, sizeof(int)));
        // a blocking (non-legacy) stream: graph capture is not allowed on the legacy default stream, and a blocking
        // stream still orders with work the caller issues on the default stream (torch events, copies)
        PLB_CUDA(cudaStreamCreate(&own_stream));
        PLB_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
        stream = own_stream; prof_stream = own_stream;
        sets[0] = GridSet{grid_in, grid_out, d_list, d_nactive};
        if (bwd_overlap && tile_scatter && sparse) {
            PLB_CUDA(cudaMalloc(&sets[1].in, gbytes));  PLB_CUDA(cudaMemset(sets[1].in, 0, gbytes));
            PLB_CUDA(cudaMalloc(&sets[1].out, gbytes)); PLB_CUDA(cudaMemset(sets[1].out, 0, gbytes));
            PLB_CUDA(cudaMalloc(&sets[1].list, n_blocks * sizeof(int)));
            PLB_CUDA(cudaMalloc(&sets[1].count, sizeof(int)));
            PLB_CUDA(cudaMemset(sets[1].count, 0, sizeof(int)));
        } else {
            bwd_overlap = false;
        }
        use_graphs = c.kernel_variant == 0 && !getenv("PLB_NO_GRAPHS");
        PLB_CUDA(cudaMalloc(&d_cursor, 4 * sizeof(int)));
        PLB_CUDA(cudaMemset(d_cursor, 0, 4 * sizeof(int)));
        stored.assign(c.max_frames, 0);
        fwd_ok.assign(c.max_frames, 0);
        svd_ok.assign(c.max_frames, 0);
        if (svd_enable && tile_scatter && sparse) {
            const size_t sb = (size_t)c.max_frames * kSvdScalars * n_pad * sizeof(T);
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            if (sb + ((size_t)4 << 30) >= free_b || cudaMalloc(&svd_store, sb) != cudaSuccess) {
                cudaGetLastError();          // not enough memory: the backward pass keeps recomputing the SVD
                svd_store = nullptr;
            }
        }
        PLB_CUDA(cudaMalloc(&store.overflow, sizeof(int)));
        PLB_CUDA(cudaMemset(store.overflow, 0, sizeof(int)));
        PLB_CUDA(cudaMalloc(&d_perm, n_pad * sizeof(int)));
        PLB_CUDA(cudaMalloc(&d_perm2, n_pad * sizeof(int)));
        k_iota<<<blocks(n_pad), kBlock>>>((int)n_pad, d_perm);
        if (const char* v = getenv("PLB_TILE")) tile_mode = atoi(v) != 0;
        if (const char* v = getenv("PLB_SVD_WARM")) svd_warm = atoi(v) != 0;
        if (const char* v = getenv("PLB_TILE_BWD")) tile_bwd = atoi(v) != 0;
        if (const char* v = getenv("PLB_SLAB_FUSED")) slab.fused = atoi(v) != 0;
        if (const char* v = getenv("PLB_SLAB_DIRECT")) slab.direct = atoi(v) != 0;
        if (const char* v = getenv("PLB_TILE_FWD_MINB")) tile_fwd_minb = atoi(v);
        if (const char* v = getenv("PLB_PDL")) pdl_enable = atoi(v) != 0;
        if (const char* v = getenv("PLB_SLAB_PUSH_INSIDE")) push_inside = atoi(v) != 0;
        if (const char* v = getenv("PLB_WINDOW_FOLLOW")) window_follow = atoi(v) != 0;
        if (const char* v = getenv("PLB_RESORT")) resort = atoi(v) != 0;
        tile_mode = tile_mode && tile_scatter && sparse && fuse;
        tile_bwd = tile_bwd && tile_mode;
        if (tile_mode) {
            chunk_cap = (c.n_particles + kChunk - 1) / kChunk + std::min(n_blocks, c.n_particles);
            PLB_CUDA(cudaMalloc(&d_chunks, (size_t)chunk_cap * sizeof(Chunk)));
            PLB_CUDA(cudaMalloc(&d_nchunks, sizeof(int)));
            chunk_grid = (c.n_particles + kChunk - 1) / kChunk;
            k_trivial_chunks<<<blocks(chunk_grid), kBlock>>>(c.n_particles, d_chunks, d_nchunks);
            if (int r = make_tile_map(&tm_out[0], sets[0].out)) return r;
            if (sets[1].out) { if (int r = make_tile_map(&tm_out[1], sets[1].out)) return r; } else tm_out[1] = tm_out[0];
            if (int r = make_tile_map(&tm_gin, g_in)) return r;
            // (without the carve-out hint the driver picked a shared-memory split that held 4 CTAs of 28.5 KB: ncu launch__occupancy_limit_shared_mem)
            const int fwd_sm = (int)chunk_smem_bytes(kFwdTiles), bwd_sm = (int)chunk_smem_bytes(kBwdTiles), carve = cudaSharedmemCarveoutMaxShared;
#define PLB_TILE_ATTR(k) do { PLB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(fwd_sm, bwd_sm))); \
                              PLB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve)); } while (0)
            PLB_TILE_ATTR((k_fwd_chunk<T, FWD_P2G, TileOcc<T>::fwd>));
            PLB_TILE_ATTR((k_fwd_chunk<T, FWD_G2P | FWD_P2G, TileOcc<T>::fwd>));
            PLB_TILE_ATTR((k_fwd_chunk<T, FWD_G2P | FWD_P2G, TileOcc<T>::fwd_hi>));
            PLB_TILE_ATTR((k_fwd_chunk<T, FWD_G2P, TileOcc<T>::fwd>));
            PLB_TILE_ATTR((k_bwd_chunk<T, BWD_G2P, TileOcc<T>::bwd_lo, false>));
            PLB_TILE_ATTR((k_bwd_chunk<T, BWD_P2G | BWD_G2P, TileOcc<T>::bwd_lo, false>));
            PLB_TILE_ATTR((k_bwd_chunk<T, BWD_P2G | BWD_G2P, TileOcc<T>::bwd_hi, true>));
            PLB_TILE_ATTR((k_bwd_chunk<T, BWD_P2G, TileOcc<T>::bwd_lo, false>));
            PLB_TILE_ATTR((k_bwd_chunk<T, BWD_P2G, TileOcc<T>::bwd_hi, true>));
#undef PLB_TILE_ATTR
        }
        PLB_CUDA(cudaDeviceSynchronize());
        return PLB_OK;
    }
    // shared memory of a chunk CTA: its scatter tiles; the TMA windows (one forward, two backward) alias them
    static size_t chunk_smem_bytes(int tiles) { return std::max((size_t)tiles * kTileVec4, (size_t)2 * kTileNodes) * sizeof(Vec4<T>); }

    // ---------------------------------------------------------------- TMA descriptors (one per grid array the particle kernels gather from)
    int make_tile_map(CUtensorMap* tm, const Vec4<T>* grid) {
        typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static Encode encode = nullptr;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            PLB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
            PLB_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
            encode = (Encode)fn;
        }
        // tensor (component, k, j, i): 4 scalars per node, node index (i n + j) n + k; box = all 4 components of 8 x 8 x 8 nodes;
        // nodes outside the grid (a window that starts at -1 or ends past n - 1) are zero-filled by the hardware
        const cuuint64_t n = (cuuint64_t)cfg.n_grid, node = 4 * sizeof(T);
        const cuuint64_t dims[4] = {4, n, n, n};
        const cuuint64_t strides[3] = {node, node * n, node * n * n};
        const cuuint32_t box[4] = {4, kTileEdge, kTileEdge, kTileEdge}, estr[4] = {1, 1, 1, 1};
        CUresult r = encode(tm, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)grid, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PLB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
        return PLB_OK;
    }

    int set_stream(void* s) override {
        // NULL / legacy stream requests keep the engine's own blocking stream (needed for graph capture)
        if (s == nullptr || (cudaStream_t)s == cudaStreamLegacy) { stream = own_stream; }
        else { stream = (cudaStream_t)s; }
        prof_stream = stream;
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        graphs.clear();
        return PLB_OK;
    }
    int synchronize() override { PLB_CUDA(cudaStreamSynchronize(stream)); return PLB_OK; }

    // frame `s` was overwritten from outside a forward substep: its stored grid and the "successor frame" links are stale
    void slot_written(int s) { stored[s] = 0; fwd_ok[s] = 0; svd_ok[s] = 0; if (s > 0) fwd_ok[s - 1] = 0; if (!slot_order.empty()) slot_order[s] = cur_order; }
    // caller-side index map of the ordering frame `slot` is stored in
    const int* perm_of_slot(int slot) const {
        if (slot_order.empty() || orders_valid == 0) return d_perm;
        const int k = slot_order[slot];
        return (k == cur_order || k < 0 || k >= orders_valid) ? d_perm : order_perm[k];
    }
    // (the permutation buffers are pooled: cudaMalloc / cudaFree per env step would serialise the host with the device)
    void clear_orders() {
        for (auto& kv : resort_q) q_spent.push_back(kv.second);
        resort_q.clear();
        cur_order = 0;
        orders_valid = 0;
        slot_order.assign(cfg.max_frames, 0);
    }
    int push_order() {          // remember the current d_perm as ordering cur_order
        if ((int)order_perm.size() <= cur_order) order_perm.resize(cur_order + 1, nullptr);
        if (!order_perm[cur_order]) PLB_CUDA(cudaMalloc(&order_perm[cur_order], n_pad * sizeof(int)));
        PLB_CUDA(cudaMemcpyAsync(order_perm[cur_order], d_perm, n_pad * sizeof(int), cudaMemcpyDeviceToDevice, stream));
        orders_valid = std::max(orders_valid, cur_order + 1);
        return PLB_OK;
    }
    bool resort_applies(int slot0) const {
        return resort && slot0 > 0 && use_graphs && sparse && tile_scatter && !slab.on && !tile_bwd && d_keys && fwd_ok[slot0 - 1] && !resort_q.count(slot0);
    }
    // re-sort frame `slot` (produced by the previous env step) in place; nothing is read back to the host
    int resort_frame(int slot) {
        const int n = cfg.n_particles;
        if (orders_valid == 0) { if (int r = push_order()) return r; }          // ordering 0 = what the last plb_sort_particles left
        k_sort_keys<T><<<blocks(n), kBlock, 0, stream>>>(P, frame_base(slot), n_pad, d_keys, d_vals);
        int bits = 1;
        while (bits < 32 && (1ull << bits) < (unsigned long long)n_blocks * 64ull) bits++;
        PLB_CUDA(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, bits, stream));
        k_permute_frame<T><<<blocks(n), kBlock, 0, stream>>>(n, n_pad, frame_base(slot), frame_tmp, d_vals2, d_perm, d_perm2);
        PLB_CUDA(cudaMemcpyAsync(frame_base(slot), frame_tmp, (size_t)24 * n_pad * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        T* mats[3] = {mat_mu, mat_lam, mat_ys};
        for (int i = 0; i < 3; i++) {
            if (!mats[i]) continue;
            k_permute_scalar<T><<<blocks(n), kBlock, 0, stream>>>(n, mats[i], frame_tmp, d_vals2);
            PLB_CUDA(cudaMemcpyAsync(mats[i], frame_tmp, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        }
        std::swap(d_perm, d_perm2);
        inv_perm_valid = false;
        int* q = nullptr;
        if (!q_spent.empty()) { q = q_spent.back(); q_spent.pop_back(); }
        else PLB_CUDA(cudaMalloc(&q, n_pad * sizeof(int)));
        PLB_CUDA(cudaMemcpyAsync(q, d_vals2, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, stream));
        resort_q[slot] = q;
        cur_order++;
        if (int r = push_order()) return r;
        slot_order[slot] = cur_order;
        launches += 3;
        if (tile_mode) {          // chunk table of the new order; its size is checked on the device against the launch grid
            PLB_CUDA(cudaMemsetAsync(d_nchunks, 0, sizeof(int), stream));
            k_build_chunks<<<blocks(n), kBlock, 0, stream>>>(n, cfg.n_grid, d_keys2, d_chunks, d_nchunks, chunk_grid, store.overflow);
            launches++;
        }
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    // after plb_step_bwd has produced the adjoint of frame `slot`: adjoint frame, frame and materials back to the previous ordering
    int unsort_frame(int slot) {
        auto it = resort_q.find(slot);
        if (it == resort_q.end()) return PLB_OK;
        const int n = cfg.n_particles;
        const size_t fb = (size_t)24 * n_pad * sizeof(T);
        k_unpermute_frame<T><<<blocks(n), kBlock, 0, stream>>>(n, n_pad, adj[cur], frame_tmp, it->second);
        PLB_CUDA(cudaMemcpyAsync(adj[cur], frame_tmp, fb, cudaMemcpyDeviceToDevice, stream));
        k_unpermute_frame<T><<<blocks(n), kBlock, 0, stream>>>(n, n_pad, frame_base(slot), frame_tmp, it->second);
        PLB_CUDA(cudaMemcpyAsync(frame_base(slot), frame_tmp, fb, cudaMemcpyDeviceToDevice, stream));
        T* mats[3] = {mat_mu, mat_lam, mat_ys};
        for (int i = 0; i < 3; i++) {
            if (!mats[i]) continue;
            k_unpermute_scalar<T><<<blocks(n), kBlock, 0, stream>>>(n, mats[i], frame_tmp, it->second);
            PLB_CUDA(cudaMemcpyAsync(mats[i], frame_tmp, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        }
        launches += 2;
        PLB_REQUIRE(cur_order > 0 && cur_order - 1 < orders_valid && order_perm[cur_order - 1], "env-step re-sort: ordering stack underflow");
        cur_order--;
        PLB_CUDA(cudaMemcpyAsync(d_perm, order_perm[cur_order], n_pad * sizeof(int), cudaMemcpyDeviceToDevice, stream));
        inv_perm_valid = false;
        slot_order[slot] = cur_order;
        q_spent.push_back(it->second);          // (back to the pool: the next user is enqueued behind the kernels that still read it)
        resort_q.erase(it);
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    int check_slot(int s) { PLB_REQUIRE(s >= 0 && s < cfg.max_frames, "frame slot out of range"); return PLB_OK; }
    int check_pf(int pf, int extra = 0) { PLB_REQUIRE(pf >= 0 && pf + extra < cfg.max_prim_frames, "primitive frame out of range"); return PLB_OK; }

    int set_materials(const double* mu, const double* lam, const double* ys) override {
        const double* src[3] = {mu, lam, ys};
        T** dst[3] = {&mat_mu, &mat_lam, &mat_ys};
        for (int i = 0; i < 3; i++) {
            if (!src[i]) continue;
            if (!*dst[i]) { PLB_CUDA(cudaMalloc(dst[i], n_pad * sizeof(T))); drop_graphs(); }     // (graphs hold the pointer by value)
            PLB_CUDA(cudaMemcpyAsync(d_stage, src[i], cfg.n_particles * sizeof(double), cudaMemcpyHostToDevice, stream));
            k_convert_perm<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(cfg.n_particles, d_stage, *dst[i], d_perm);
            launches++;
            PLB_CUDA(cudaStreamSynchronize(stream));
        }
        return PLB_OK;
    }

    int upload_aos(const double* x, const double* v, const double* F, const double* C, double** dx, double** dv, double** dF, double** dC) {
        size_t n = cfg.n_particles;
        *dx = x ? d_stage : nullptr; *dv = v ? d_stage + 3 * n : nullptr;
        *dF = F ? d_stage + 6 * n : nullptr; *dC = C ? d_stage + 15 * n : nullptr;
        if (x) PLB_CUDA(cudaMemcpyAsync(*dx, x, 3 * n * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (v) PLB_CUDA(cudaMemcpyAsync(*dv, v, 3 * n * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (F) PLB_CUDA(cudaMemcpyAsync(*dF, F, 9 * n * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (C) PLB_CUDA(cudaMemcpyAsync(*dC, C, 9 * n * sizeof(double), cudaMemcpyHostToDevice, stream));
        return PLB_OK;
    }
    int download_aos(T* frame, double* x, double* v, double* F, double* C, const int* perm = nullptr) {
        size_t n = cfg.n_particles;
        double *dx = x ? d_stage : nullptr, *dv = v ? d_stage + 3 * n : nullptr, *dF = F ? d_stage + 6 * n : nullptr,
               *dC = C ? d_stage + 15 * n : nullptr;
        k_unpack_frame<T><<<blocks(n), kBlock, 0, stream>>>((int)n, n_pad, frame, perm ? perm : d_perm, dx, dv, dF, dC);
        launches++;
        if (x) PLB_CUDA(cudaMemcpyAsync(x, dx, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (v) PLB_CUDA(cudaMemcpyAsync(v, dv, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (F) PLB_CUDA(cudaMemcpyAsync(F, dF, 9 * n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (C) PLB_CUDA(cudaMemcpyAsync(C, dC, 9 * n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        PLB_CUDA(cudaStreamSynchronize(stream));
        return PLB_OK;
    }

    int set_frame(int slot, const double* x, const double* v, const double* F, const double* C) override {
        if (int r = check_slot(slot)) return r;
        slot_written(slot);
        double *dx, *dv, *dF, *dC;
        if (int r = upload_aos(x, v, F, C, &dx, &dv, &dF, &dC)) return r;
        k_pack_frame<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(cfg.n_particles, n_pad, frame_base(slot), d_perm, dx, dv, dF, dC);
        launches++;
        PLB_CUDA(cudaStreamSynchronize(stream));      // the host arrays may be reused by the caller
        return PLB_OK;
    }
    int get_frame(int slot, double* x, double* v, double* F, double* C) override {
        if (int r = check_slot(slot)) return r;
        return download_aos(frame_base(slot), x, v, F, C, perm_of_slot(slot));
    }
    int copy_frame(int src, int dst) override {
        if (int r = check_slot(src)) return r;
        if (int r = check_slot(dst)) return r;
        if (src == dst) return PLB_OK;
        slot_written(dst);
        if (!slot_order.empty()) slot_order[dst] = slot_order[src];          // (a copy keeps the ordering it was stored in)
        PLB_CUDA(cudaMemcpyAsync(frame_base(dst), frame_base(src), (size_t)24 * n_pad * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        return PLB_OK;
    }
    // Re-order the particles of `slot` by (4^3 block, cell) of their base cell so that a warp's 32 particles share
    // stencil nodes (coalesced gathers, few distinct scatter addresses).  All other slots become stale; the host-side
    // order is kept through d_perm.  Called when a new state is installed (reset / set_state of frame 0).
    int sort_particles(int slot) override {
        if (int r = check_slot(slot)) return r;
        const int n = cfg.n_particles;
        if (!d_keys) {
            PLB_CUDA(cudaMalloc(&d_keys, n_pad * sizeof(unsigned)));  PLB_CUDA(cudaMalloc(&d_keys2, n_pad * sizeof(unsigned)));
            PLB_CUDA(cudaMalloc(&d_vals, n_pad * sizeof(int)));       PLB_CUDA(cudaMalloc(&d_vals2, n_pad * sizeof(int)));
            PLB_CUDA(cudaMalloc(&frame_tmp, (size_t)24 * n_pad * sizeof(T)));
            cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 32, stream);
            PLB_CUDA(cudaMalloc(&d_cub, cub_bytes));
        }
        k_sort_keys<T><<<blocks(n), kBlock, 0, stream>>>(P, frame_base(slot), n_pad, d_keys, d_vals);
        int bits = 1;
        while (bits < 32 && (1ull << bits) < (unsigned long long)n_blocks * 64ull) bits++;      // ceil(log2(largest key + 1))
        PLB_CUDA(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, bits, stream));
        k_permute_frame<T><<<blocks(n), kBlock, 0, stream>>>(n, n_pad, frame_base(slot), frame_tmp, d_vals2, d_perm, d_perm2);
        PLB_CUDA(cudaMemcpyAsync(frame_base(slot), frame_tmp, (size_t)24 * n_pad * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        T* mats[3] = {mat_mu, mat_lam, mat_ys};
        for (int i = 0; i < 3; i++) {
            if (!mats[i]) continue;
            k_permute_scalar<T><<<blocks(n), kBlock, 0, stream>>>(n, mats[i], frame_tmp, d_vals2);
            PLB_CUDA(cudaMemcpyAsync(mats[i], frame_tmp, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        }
        std::swap(d_perm, d_perm2);
        inv_perm_valid = false;
        clear_orders();
        launches += 3;
        if (tile_mode) {
            PLB_CUDA(cudaMemsetAsync(d_nchunks, 0, sizeof(int), stream));
            k_build_chunks<<<blocks(n), kBlock, 0, stream>>>(n, cfg.n_grid, d_keys2, d_chunks, d_nchunks, chunk_cap, store.overflow);
            launches++;
            int nch = 0;
            PLB_CUDA(cudaMemcpyAsync(&nch, d_nchunks, sizeof(int), cudaMemcpyDeviceToHost, stream));
            PLB_CUDA(cudaStreamSynchronize(stream));
            PLB_REQUIRE(nch > 0 && nch <= chunk_cap, "chunk table overflow");
            // (captured graphs hold the launch geometry; with the env-step re-sort the table is rebuilt without a read-back, so
            //  the grid keeps a margin for the block count growing as the body spreads)
            const int want_grid = std::min(chunk_cap, resort ? nch + nch / 4 + 256 : nch + nch / 16 + 64);
            if (nch > chunk_grid || (resort && chunk_grid < want_grid)) {
                chunk_grid = want_grid;
                drop_graphs();
            }
        }
        PLB_CUDA(cudaGetLastError());
        std::fill(stored.begin(), stored.end(), 0);
        std::fill(fwd_ok.begin(), fwd_ok.end(), 0);
        std::fill(svd_ok.begin(), svd_ok.end(), 0);
        // size the forward-grid store from the active-block count of this frame (2x margin + 256 blocks)
        if (sparse && cfg.kernel_variant == 0) {
            k_mark_only<T><<<blocks(n), kBlock, 0, stream>>>(P, frame_base(slot), n_pad, d_flags);
            if (env_list) {          // env-step lists are dilated by one block: size the store for the dilated count
                cudaMemsetAsync(d_nactive, 0, sizeof(int), stream);
                k_dilate_flags<<<(n_blocks + 255) / 256, 256, 0, stream>>>(cfg.n_grid / 4, d_flags, d_flags2);
                k_compact_mark<<<(n_blocks + 255) / 256, 256, 0, stream>>>(n_blocks, d_flags2, d_list, d_nactive, d_listed);
                launches += 2;
            } else {
                compact_blocks();
            }
            int cnt = 0;
            PLB_CUDA(cudaMemcpyAsync(&cnt, d_nactive, sizeof(int), cudaMemcpyDeviceToHost, stream));
            PLB_CUDA(cudaStreamSynchronize(stream));
            listed_est = cnt;
            int want = std::min(n_blocks, (env_list ? cnt + cnt / 2 : 2 * cnt) + 256);
            if (want > store.cap) {
                cudaFree(store.vals); cudaFree(store.ids); cudaFree(store.cnt);
                store.vals = nullptr; store.ids = nullptr; store.cnt = nullptr; store.cap = 0;
                size_t vb = (size_t)cfg.max_frames * want * kBlkNodes * sizeof(Vec4<T>);
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                if (vb + ((size_t)2 << 30) < free_b && cudaMalloc(&store.vals, vb) == cudaSuccess) {
                    PLB_CUDA(cudaMalloc(&store.ids, (size_t)cfg.max_frames * want * sizeof(int)));
                    PLB_CUDA(cudaMalloc(&store.cnt, (size_t)cfg.max_frames * sizeof(int)));
                    store.cap = want;
                } else {
                    cudaGetLastError();      // not enough memory: keep recomputing P2G in the backward pass
                    store.vals = nullptr;
                }
                for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
                graphs.clear();
            }
        }
        return PLB_OK;
    }
    int frame_ptr(int slot, void** ptr, long long* np, int* sb) override {
        if (int r = check_slot(slot)) return r;
        *ptr = frame_base(slot); *np = n_pad; *sb = (int)sizeof(T);
        return PLB_OK;
    }

    // ---------------------------------------------------------------- primitives
    double* pose(int pf, int k) { return traj.data() + ((size_t)pf * PLB_MAX_PRIM + k) * 8; }
    double* velo(int pf, int k) { return vel.data() + ((size_t)pf * PLB_MAX_PRIM + k) * 8; }
    int upload_poses(int pf0, int n) {
        size_t off = (size_t)pf0 * PLB_MAX_PRIM * 8;
        PLB_CUDA(cudaMemcpyAsync(d_traj + off, traj.data() + off, (size_t)n * PLB_MAX_PRIM * 8 * sizeof(double),
                                 cudaMemcpyHostToDevice, stream));
        return PLB_OK;
    }
    int set_prim_state(int pf, int k, const double* s) override {
        if (int r = check_pf(pf)) return r;
        PLB_REQUIRE(k >= 0 && k < cfg.n_primitives, "primitive index");
        std::memcpy(pose(pf, k), s, 8 * sizeof(double));
        return upload_poses(pf, 1);
    }
    int get_prim_state(int pf, int k, double* s) override {
        if (int r = check_pf(pf)) return r;
        PLB_REQUIRE(k >= 0 && k < cfg.n_primitives, "primitive index");
        std::memcpy(s, pose(pf, k), 8 * sizeof(double));
        return PLB_OK;
    }
    int copy_prim_frame(int src, int dst) override {
        if (int r = check_pf(src)) return r;
        if (int r = check_pf(dst)) return r;
        std::memcpy(pose(dst, 0), pose(src, 0), PLB_MAX_PRIM * 8 * sizeof(double));
        return upload_poses(dst, 1);
    }
    void drop_graphs() {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        graphs.clear();
    }
    // kernel arguments (PrimSet with the softness, the material pointers) are baked into captured graphs by value:
    // a change must drop the cache, or a replayed forward graph would run with the old value beside a freshly captured
    // backward graph with the new one
    int set_softness(double s) override {
        bool changed = false;
        for (int k = 0; k < cfg.n_primitives; k++) { changed = changed || prims.s[k].softness != (T)s; prims.s[k].softness = (T)s; }
        if (changed) drop_graphs();
        return PLB_OK;
    }
    int set_action(int step, int S, const double* a, int n) override {
        PLB_REQUIRE(n == action_total, "action length != sum of action dims");
        PLB_REQUIRE(step >= 0 && step < max_steps_actions && S > 0, "bad step");
        if (int r = check_pf((step + 1) * S - 1)) return r;
        for (int i = 0; i < n; i++) actions[(size_t)step * action_total + i] = std::min(1.0, std::max(-1.0, a[i]));
        for (int k = 0; k < cfg.n_primitives; k++) {
            const kin::Desc& d = kdesc[k];
            if (d.action_dim == 0) continue;
            const double* ak = &actions[(size_t)step * action_total + action_off[k]];
            for (int f = step * S; f < (step + 1) * S; f++) {
                double* vv = velo(f, k);
                for (int i = 0; i < 3; i++) vv[i] = ak[i] * d.action_scale[i] / S;
                if (d.action_dim > 3) for (int i = 0; i < 3; i++) vv[3 + i] = ak[3 + i] * d.action_scale[3 + i] / S;
                if (d.type == PRIM_CHOPSTICKS) vv[6] = ak[6] * d.action_scale[6] / S;
            }
        }
        return PLB_OK;
    }
    int kinematics(int pf, int n) override {
        if (int r = check_pf(pf, n)) return r;
        for (int f = pf; f < pf + n; f++)
            for (int k = 0; k < cfg.n_primitives; k++) {
                const double* vv = velo(f, k);
                kin::fk_forward(kdesc[k], pose(f, k), vv, vv + 3, vv[6], pose(f + 1, k));
            }
        return upload_poses(pf + 1, n);
    }

    // ---------------------------------------------------------------- substeps
    // CTAs of the list-driven grid kernels (2 blocks of 64 nodes per CTA and round).  These kernels are latency chains over a few
    // hundred to a few thousand blocks: n_blocks / 2 CTAs (1184 at 128^3) made the 96-register grid adjoint run in two waves
    // with the SMs idle half of its 14 us (ncu at move100k, profiles/r2_grid_kernels_ncu_summary.md), so the launch is sized from
    // the block count seen at the last sort (1.25x margin, whole multiples of the 148 SMs) and capped at one wave.
    int listed_est = 0;
    int sparse_ctas(int wave_cap = 148 * 8) const {
        if (listed_est <= 0) return std::min(std::max(n_blocks / 2, 1), wave_cap);
        const int want = (listed_est + listed_est / 4 + 1) / 2;
        return std::min(std::max((want + 147) / 148 * 148, 148), wave_cap);
    }
    void compact_blocks() {
        cudaMemsetAsync(d_nactive, 0, sizeof(int), stream);
        k_compact<<<(n_blocks + 255) / 256, 256, 0, stream>>>(n_blocks, d_flags, d_list, d_nactive);
        launches++;
    }
    // Enqueue one forward substep.  Slots/poses are given as SlotRef so the same code serves direct launches
    // (absolute indices) and graph capture (cursor-relative indices).
    // grid stage of a forward substep: (halo) + active-block list + grid operator (+ store of the forward grid for slot `si`)
    // forward graphs in env-list mode: single GPU, forward-grid store present
    bool slab_fused() const { return slab.peer_ready && slab.fused && env_list && store.vals; }
    bool env_list_mode() const { return env_list && sparse && tile_scatter && store.vals && (!slab.on || slab_fused()); }
    // mode: 0 none, 1 fused receive (wait + add the inbox data), 2 direct halo (wait only), 3 send + receive inside the grid kernel
    bool push_inside = true;        // PLB_SLAB_PUSH_INSIDE=0: separate k_halo_push2 launch ahead of the grid kernel (mode 1)
    // A grid kernel that sends AND waits (mode 3) must have all its CTAs co-resident: a CTA spinning on the neighbour's flag
    // would otherwise hold the slot of a CTA of its own kernel that has not pushed yet, and the neighbour -- waiting for that push --
    // does the same (seen on 2 GPUs with the float64 kernels, whose 194 registers allow 2 CTAs per SM where the launch assumed 5).
    // The launch is therefore capped by the occupancy the runtime reports for the instantiation.
    int occ_grid_fwd = 0, occ_grid_bwd = 0;
    int coresident_cap(bool bwd) {
        int& occ = bwd ? occ_grid_bwd : occ_grid_fwd;
        if (occ == 0) {
            int o = 0;
            cudaError_t e = bwd ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_grid_bwd_sparse_v2<T>, kBlock, 0)
                                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_grid_fwd_sparse<T>, kBlock, 0);
            if (e != cudaSuccess) { cudaGetLastError(); o = 1; }
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
            occ = std::max(1, o) * std::max(1, sms);
        }
        return occ;
    }
    HaloIn halo_in(int mode) const {
        HaloIn h;
        for (int side = 0; side < 2; side++) { h.inbox[side] = (mode && slab.has[side]) ? slab.inbox[side] : nullptr; h.g[side] = slab.geom[side]; }
        h.seq = slab.seq; h.err = slab.err; h.on = mode;
        for (int side = 0; side < 2; side++) { h.peer[side] = (mode == 3 && slab.has[side]) ? slab.peer[side] : nullptr; h.pg[side] = slab.geom[side]; }
        h.seq_w = slab.seq; h.done = slab.done;
        return h;
    }
    bool slab_direct() const {
        if (!(slab_fused() && slab.direct && tile_mode && !tile_bwd && grid_bwd_v2 && sets[1].in && slab.g_out2)) return false;
        for (int side = 0; side < 2; side++)
            if (slab.has[side]) for (int w = 0; w < 4; w++) if (!slab.peer_grid[side][w]) return false;
        return true;
    }
    // adjoint-of-grid_out buffer the backward scatter of substep j goes into (and its grid adjoint consumes): g_out, or with the
    // direct halo the parity-j buffer whose copies on the neighbours receive the zone contributions
    Vec4<T>* cur_gout = nullptr; int cur_gout_which = -1;
    void select_gout(int j, bool direct) {
        if (direct) { cur_gout_which = 2 + (j & 1); cur_gout = local_grid(cur_gout_which); }
        else { cur_gout_which = -1; cur_gout = g_out; }
    }
    PeerHalo<Vec4<T>> gout_peers() const { return cur_gout_which >= 0 ? peer_halo(cur_gout_which) : no_peers<Vec4<T>>(); }
    HaloOut gout_publish() const { return cur_gout_which >= 0 ? halo_out() : halo_out_none(); }
    Vec4<T>* local_grid(int which) const { return which == 0 ? sets[0].in : which == 1 ? sets[1].in : which == 2 ? g_out : slab.g_out2; }
    // which: 0/1 grid_in of even/odd substeps, 2/3 g_out of even/odd substeps
    PeerHalo<Vec4<T>> peer_halo(int which) const {
        PeerHalo<Vec4<T>> p = no_peers<Vec4<T>>();
        for (int side = 0; side < 2; side++)
            if (slab.has[side]) { p.grid[side] = slab.peer_grid[side][which]; p.lo[side] = slab.zlo[side]; p.hi[side] = slab.zhi[side]; }
        return p;
    }
    HaloOut halo_out() const {
        HaloOut h;
        for (int side = 0; side < 2; side++) h.peer[side] = slab.has[side] ? slab.peer[side] : nullptr;
        h.seq = slab.seq; h.done = slab.done;
        return h;
    }
    // one launch: my listed zone blocks of `grid` -> both neighbours' inboxes, published by the last CTA (k_halo_push2)
    void halo_push_fused(const Vec4<T>* grid, const int* list, const int* count) {
        k_halo_push2<T><<<148, kBlock, 0, stream>>>(cfg.n_grid, grid, list, count, slab.has[0] ? slab.peer[0] : nullptr, slab.has[1] ? slab.peer[1] : nullptr,
                                                    slab.geom[0], slab.geom[1], slab.seq, slab.done);
        launches++;
    }
    // list of the env step that starts at frame `slot`: blocks touched by that frame, dilated by one block
    void enqueue_env_list(SlotRef slot) {
        const int nbx = cfg.n_grid / 4, nb = (n_blocks + 255) / 256;
        cudaMemsetAsync(d_nactive, 0, sizeof(int), stream);
        k_mark_slot<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, slot, d_flags);
        if (slab_direct()) {
            // the neighbours add into my zone planes directly; what they put into blocks outside my list (the far edge of a zone,
            // which only their material reaches) is never consumed or cleared by my grid kernels: wipe the zones once per env step,
            // before my flags go out (the neighbours scatter again only after they have received them)
            const size_t plane = (size_t)cfg.n_grid * cfg.n_grid * sizeof(Vec4<T>);
            for (int side = 0; side < 2; side++)
                if (slab.has[side])
                    for (int w = 0; w < 4; w++)
                        cudaMemsetAsync((char*)local_grid(w) + (size_t)slab.zlo[side] * plane, 0, (size_t)(slab.zhi[side] - slab.zlo[side]) * plane, stream);
        }
        if (slab_fused()) {
            // the neighbours' particles scatter into the zones too: exchange the zone flags once per env step, so that both sides
            // list (and push / receive) the same zone blocks for all of its substeps
            k_halo_push_flags<<<32, 256, 0, stream>>>(cfg.n_grid, d_flags, slab.has[0] ? slab.peer[0] : nullptr, slab.has[1] ? slab.peer[1] : nullptr,
                                                      slab.geom[0], slab.geom[1], slab.seq, slab.done);
            k_halo_or_flags<<<32, 256, 0, stream>>>(cfg.n_grid, d_flags, halo_in(1));
            launches += 2;
        }
        k_dilate_flags<<<nb, 256, 0, stream>>>(nbx, d_flags, d_flags2);
        k_compact_mark<<<nb, 256, 0, stream>>>(n_blocks, d_flags2, d_list, d_nactive, d_listed);
        launches += 3;
    }
    // the frame the env step produced must still lie inside its list (else *store.overflow = 2 -> check_overflow reports it)
    void enqueue_env_list_check(SlotRef slot) {
        k_mark_slot<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, slot, d_flags);
        k_check_listed<<<(n_blocks + 255) / 256, 256, 0, stream>>>(n_blocks, d_flags, d_listed, store.overflow);
        launches += 2;
        if (slab.on && slab.err) {       // ownership is fixed for the episode (no migration): the material must stay within owned planes + halo
            k_check_margin<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, slot, slab.has[0] ? slab.own_lo - slab.w : 0,
                                                                          slab.has[1] ? slab.own_hi + slab.w : cfg.n_grid, slab.err);
            launches++;
        }
    }
    // gin: the buffer the substep's scatter went into (direct halo alternates between the two grid sets' by substep parity)
    void enqueue_grid_fwd_stage(SlotRef si, SlotRef pf, bool fixed_list = false, Vec4<T>* gin = nullptr) {
        if (!gin) gin = grid_in;
        prof_begin(K_GRID_FWD);
        if (sparse) {
            if (fixed_list) {
                // (the list in d_list / d_nactive was built by enqueue_env_list for the whole env step)
                if (slab_fused()) {
                    const bool direct = slab_direct();
                    if (!direct && !push_inside) halo_push_fused(gin, d_list, d_nactive);
                    launch_k(pdl_on(), k_grid_fwd_sparse<T>, sparse_ctas(std::min(148 * 8, coresident_cap(false))), kBlock, 0, stream, P, prims, d_traj, pf, gin, grid_out, 1, d_list, d_nactive, store, si, halo_in(direct ? 2 : push_inside ? 3 : 1));
                    prof_end();
                    launches++;
                    return;
                }
            } else if (slab.peer_ready) {
                k_halo_next_seq<<<1, 1, 0, stream>>>(slab.seq);
                cudaMemsetAsync(d_nactive, 0, sizeof(int), stream);
                k_compact_stamped<<<(n_blocks + 255) / 256, 256, 0, stream>>>(n_blocks, d_flags, d_list, d_nactive, slab.listed_stamp, slab.seq);
                halo_exchange(grid_in);
                for (int side = 0; side < 2; side++)
                    if (slab.has[side])
                        k_halo_append<<<32, 256, 0, stream>>>(cfg.n_grid, slab.inbox[side], slab.geom[side], slab.own_lo, slab.own_hi, slab.seq,
                                                              slab.listed_stamp, d_list, d_nactive);
                halo_add_inbox(grid_in);
                launches += 8;
            } else {
                compact_blocks();
            }
            launch_k(fixed_list && pdl_on(), k_grid_fwd_sparse<T>, sparse_ctas(), kBlock, 0, stream, P, prims, d_traj, pf, grid_in, grid_out, 1, d_list, d_nactive, store, si, halo_in(0));
        } else {
            k_grid_fwd<T><<<blocks(n_nodes), kBlock, 0, stream>>>(P, prims, d_traj, pf, grid_in, grid_out, 1, n_nodes);
        }
        prof_end();
        launches++;
    }
    void enqueue_fwd(SlotRef si, SlotRef so, SlotRef pf) {
        const int nb = blocks(cfg.n_particles);
        prof_begin(K_P2G);
        launch_p2g(si, so, 1);
        prof_end();
        enqueue_grid_fwd_stage(si, pf);
        prof_begin(K_G2P);
        k_g2p<T><<<nb, kBlock, 0, stream>>>(P, frames, n_pad, si, so, grid_out);
        prof_end();
        launches += 2;
    }
    // n >= 2 forward substeps with G2P(i-1) and P2G(i) fused into one particle kernel (refs are cursor-relative or absolute)
    void set_pdl_inhibit() { pdl_inhibit = slab.on && !(slab_fused() && push_inside && !slab_direct()); }
    void enqueue_fwd_fused(int n, SlotRef (*mk)(const Engine*, int, int), bool fixed_list = false) {
        set_pdl_inhibit();
        const int nb = blocks(cfg.n_particles);
        unsigned char* fl = (sparse && !fixed_list) ? d_flags : nullptr;
        if (fixed_list) enqueue_env_list(mk(this, 0, 0));
        if (tile_mode && window_follow) {          // TMA windows follow the material (k_chunk_origins)
            k_chunk_origins<T><<<chunk_grid, kBlock, 0, stream>>>(P, frames, n_pad, mk(this, 0, 0), d_chunks, d_nchunks);
            launches++;
        }
        if (tile_mode) {
            // chunked kernels: one CTA per <= 128 particles of one grid block, grid_out window by TMA (plb_tile.cuh)
            const size_t sm = chunk_smem_bytes(kFwdTiles);
            const SlotRef none = abs_ref(0);
            const bool direct = fixed_list && slab_direct();          // scatter of substep i -> grid set i & 1, mirrored into the neighbours' copies
            const HaloOut ho = direct ? halo_out() : halo_out_none();
            auto gin = [&](int i) { return direct ? sets[i & 1].in : grid_in; };
            auto ph = [&](int i) { return direct ? peer_halo(i & 1) : no_peers<Vec4<T>>(); };
            prof_begin(K_P2G);
            k_fwd_chunk<T, FWD_P2G, TileOcc<T>::fwd><<<chunk_grid, kBlock, sm, stream>>>(tm_out[0], P, frames, n_pad, mk(this, 0, 0), none, mk(this, 1, 0), material(),
                                                                                      chunk_table(), grid_out, gin(0), fl, svd_store, 0, ph(0), ho);
            prof_end();
            enqueue_grid_fwd_stage(mk(this, 0, 0), mk(this, 2, 0), fixed_list, gin(0));
            for (int i = 1; i < n; i++) {
                prof_begin(K_G2P_P2G);
                auto kern = tile_fwd_minb >= 6 ? k_fwd_chunk<T, FWD_G2P | FWD_P2G, TileOcc<T>::fwd_hi> : k_fwd_chunk<T, FWD_G2P | FWD_P2G, TileOcc<T>::fwd>;
                launch_k(fixed_list && pdl_on(), kern, chunk_grid, kBlock, sm, stream, tm_out[0], P, frames, n_pad, mk(this, 0, i - 1), mk(this, 0, i), mk(this, 1, i), material(),
                         chunk_table(), grid_out, gin(i), fl, svd_store, svd_warm ? 1 : 0, ph(i), ho);
                prof_end();
                enqueue_grid_fwd_stage(mk(this, 0, i), mk(this, 2, i), fixed_list, gin(i));
                launches++;
            }
            prof_begin(K_G2P);
            launch_k(fixed_list && pdl_on(), k_fwd_chunk<T, FWD_G2P, TileOcc<T>::fwd>, chunk_grid, kBlock, sm, stream, tm_out[0], P, frames, n_pad, mk(this, 0, n - 1), none,
                     mk(this, 1, n - 1), material(), chunk_table(), grid_out, grid_in, (unsigned char*)nullptr, (T*)nullptr, 0, no_peers<Vec4<T>>(), halo_out_none());
            prof_end();
            launches += 2;
            if (fixed_list) enqueue_env_list_check(mk(this, 1, n - 1));
            return;
        }
        prof_begin(K_P2G);
        launch_p2g(mk(this, 0, 0), mk(this, 1, 0), 1, !fixed_list);
        prof_end();
        enqueue_grid_fwd_stage(mk(this, 0, 0), mk(this, 2, 0), fixed_list);
        const int nbc = blocks(cfg.n_particles, cta);
        const size_t sm = tile_smem_bytes(cta);
        for (int i = 1; i < n; i++) {
            prof_begin(K_G2P_P2G);
            auto kern = fwd_minb >= 6 ? k_g2p_p2g_warp<T, OccSel<T>::fwd_hi> : k_g2p_p2g_warp<T, OccSel<T>::fwd_lo>;
            kern<<<nbc, cta, sm, stream>>>(P, frames, n_pad, mk(this, 0, i - 1), mk(this, 0, i), mk(this, 1, i), material(), grid_out, grid_in, fl, flush_mode, svd_store);
            prof_end();
            enqueue_grid_fwd_stage(mk(this, 0, i), mk(this, 2, i), fixed_list);
            launches++;
        }
        prof_begin(K_G2P);
        k_g2p<T><<<nb, kBlock, 0, stream>>>(P, frames, n_pad, mk(this, 0, n - 1), mk(this, 1, n - 1), grid_out);
        prof_end();
        launches += 2;
        if (fixed_list) enqueue_env_list_check(mk(this, 1, n - 1));
    }
    // backward stages of one substep, on grid set `gs`, enqueued on stream `st`
    void enqueue_bwd_grid_pre(SlotRef si, SlotRef pf, bool restore, const GridSet& gs, cudaStream_t st) {   // forward grid of the substep + grid_out
        const int ng = blocks(n_nodes);
        GridStore<T> nostore{nullptr, nullptr, nullptr, nullptr, 0};
        prof_begin(K_P2G_RECOMPUTE, st);
        if (restore) {
            k_restore_blocks<T><<<sparse_ctas(), kBlock, 0, st>>>(cfg.n_grid, gs.in, gs.list, gs.count, store, si);
        } else {
            launch_p2g(si, si, 0);                      // (set 0 on the main stream: the non-stored path is never overlapped)
            if (sparse) compact_blocks();
        }
        prof_end(); prof_begin(K_GRID_FWD_RECOMPUTE, st);
        if (sparse)
            k_grid_fwd_sparse<T><<<sparse_ctas(), kBlock, 0, st>>>(P, prims, d_traj, pf, gs.in, gs.out, 0, gs.list, gs.count, nostore, si, halo_in(0));
        else
            k_grid_fwd<T><<<ng, kBlock, 0, st>>>(P, prims, d_traj, pf, gs.in, gs.out, 0, n_nodes);
        prof_end();
        launches += 2;
    }
    bool pdl_bwd = false;          // set by enqueue_bwd_fused around the launches that follow a kernel with pdl_launch() in the stream
    void enqueue_bwd_grid_adj(SlotRef pf, const GridSet& gs) {               // (halo of g_out) + grid_op.grad
        const int ng = blocks(n_nodes);
        prof_begin(K_GRID_BWD);
        Vec4<T>* g_out = cur_gout ? cur_gout : this->g_out;
        if (slab_fused() && sparse && grid_bwd_v2) {
            const bool direct = cur_gout_which >= 0;
            if (!direct && !push_inside) halo_push_fused(g_out, gs.list, gs.count);
            launch_k(pdl_bwd && pdl_on(), k_grid_bwd_sparse_v2<T>, sparse_ctas(std::min(148 * 5, coresident_cap(true))), kBlock, 0, stream, P, prims, d_traj, pf, gs.in, g_out, g_in, 1, d_prim_grad, gs.list, gs.count, own_lo(), own_hi(), halo_in(direct ? 2 : push_inside ? 3 : 1));
            prof_end();
            launches++;
            return;
        }
        if (slab.peer_ready) {
            k_halo_next_seq<<<1, 1, 0, stream>>>(slab.seq);
            halo_exchange(g_out);
            halo_add_inbox(g_out);
            launches += 6;
        }
        if (sparse && grid_bwd_v2 && !slab.on)         // (slab runs keep the array form: the register form was validated on one GPU only)
            launch_k(pdl_bwd && pdl_on(), k_grid_bwd_sparse_v2<T>, sparse_ctas(148 * 5), kBlock, 0, stream, P, prims, d_traj, pf, gs.in, g_out, g_in, 1, d_prim_grad, gs.list, gs.count, own_lo(), own_hi(), halo_in(0));
        else if (sparse)
            k_grid_bwd_sparse<T><<<sparse_ctas(), kBlock, 0, stream>>>(P, prims, d_traj, pf, gs.in, g_out, g_in, 1, d_prim_grad, gs.list, gs.count, own_lo(), own_hi());
        else
            k_grid_bwd<T><<<ng, kBlock, 0, stream>>>(P, prims, d_traj, pf, gs.in, g_out, g_in, 1, d_prim_grad, n_nodes);
        prof_end();
        launches++;
    }
    // (grid set -> its TMA descriptor)
    const CUtensorMap& tm_of(const GridSet& gs) const { return gs.out == sets[1].out ? tm_out[1] : tm_out[0]; }
    // s_next: the successor frame of si (only dereferenced when next_ok)
    void launch_g2p_bwd(SlotRef si, T* a_next, T* a_cur, const GridSet& gs, bool next_ok, bool chunked = false, SlotRef s_next = SlotRef{nullptr, 0, 0}) {
        Vec4<T>* g_out = cur_gout ? cur_gout : this->g_out;
        prof_begin(K_G2P_BWD);
        if (chunked) {
            k_bwd_chunk<T, BWD_G2P, TileOcc<T>::bwd_lo, false><<<chunk_grid, kBlock, chunk_smem_bytes(kBwdTiles), stream>>>(
                tm_gin, tm_of(gs), P, frames, n_pad, s_next, si, next_ok ? 1 : 0, a_next, a_cur, material(), chunk_table(), g_in, gs.out, g_out, (T*)nullptr);
        } else if (tile_scatter) {
            const int nbc = blocks(cfg.n_particles, cta);
            const size_t sm = tile_smem_bytes(cta);
            k_g2p_bwd_warp<T><<<nbc, cta, sm, stream>>>(P, frames, n_pad, si, next_ok ? 1 : 0, a_next, a_cur, gs.out, g_out, flush_mode, gout_peers(), gout_publish());
        } else {
            k_g2p_bwd<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, si, a_next, a_cur, gs.out, g_out);
        }
        prof_end();
        launches++;
    }
    void launch_p2g_bwd(SlotRef si, T* a_next, T* a_cur, bool svd, bool chunked = false) {
        prof_begin(K_P2G_BWD);
        if (chunked) {
            const size_t sm = chunk_smem_bytes(kBwdTiles);
            if (svd && bwd_minb >= 4)
                k_bwd_chunk<T, BWD_P2G, TileOcc<T>::bwd_hi, true><<<chunk_grid, kBlock, sm, stream>>>(tm_gin, tm_out[0], P, frames, n_pad, si, si, 0, a_next, a_cur, material(),
                                                                                                   chunk_table(), g_in, grid_out, g_out, svd_store);
            else
                k_bwd_chunk<T, BWD_P2G, TileOcc<T>::bwd_lo, false><<<chunk_grid, kBlock, sm, stream>>>(tm_gin, tm_out[0], P, frames, n_pad, si, si, 0, a_next, a_cur, material(),
                                                                                                    chunk_table(), g_in, grid_out, g_out, (T*)nullptr);
        } else if (svd) k_p2g_bwd<T, true><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, si, a_next, a_cur, material(), g_in, svd_store);
        else k_p2g_bwd<T, false><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, si, a_next, a_cur, material(), g_in, (T*)nullptr);
        prof_end();
        launches++;
    }
    void launch_bwd_fused(SlotRef s_cur, SlotRef s_prev, T* a_next, T* a_cur, const GridSet& gs, bool svd) {
        const int nbc = blocks(cfg.n_particles, cta);
        const size_t sm = tile_smem_bytes(cta);
        Vec4<T>* g_out = cur_gout ? cur_gout : this->g_out;
        prof_begin(K_P2G_BWD_G2P_BWD);
        if (tile_bwd) {
            if (svd && bwd_minb >= 4)
                k_bwd_chunk<T, BWD_P2G | BWD_G2P, TileOcc<T>::bwd_hi, true><<<chunk_grid, kBlock, chunk_smem_bytes(kBwdTiles), stream>>>(
                    tm_gin, tm_of(gs), P, frames, n_pad, s_cur, s_prev, 1, a_next, a_cur, material(), chunk_table(), g_in, gs.out, g_out, svd_store);
            else
                k_bwd_chunk<T, BWD_P2G | BWD_G2P, TileOcc<T>::bwd_lo, false><<<chunk_grid, kBlock, chunk_smem_bytes(kBwdTiles), stream>>>(
                    tm_gin, tm_of(gs), P, frames, n_pad, s_cur, s_prev, 1, a_next, a_cur, material(), chunk_table(), g_in, gs.out, g_out, (T*)nullptr);
            prof_end();
            launches++;
            return;
        }
        auto kern = svd ? (bwd_minb >= 4 ? k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_hi, true> : k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_lo, true>)
                        : (bwd_minb >= 4 ? k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_hi, false> : k_p2g_bwd_g2p_bwd_warp<T, OccSel<T>::bwd_lo, false>);
        launch_k(pdl_bwd && pdl_on(), kern, nbc, cta, sm, stream, P, frames, n_pad, s_cur, s_prev, a_next, a_cur, material(), g_in, gs.out, g_out, flush_mode, svd_store,
                 gout_peers(), gout_publish());
        prof_end();
        launches++;
    }
    // One backward substep; `restore` = the forward grid of this slot is in the store; next_ok = slot si+1 holds G2P's output.
    void enqueue_bwd(SlotRef si, SlotRef pf, bool restore, bool next_ok, bool svd, T* a_next, T* a_cur) {
        enqueue_bwd_grid_pre(si, pf, restore, sets[0], stream);
        launch_g2p_bwd(si, a_next, a_cur, sets[0], next_ok);
        enqueue_bwd_grid_adj(pf, sets[0]);
        launch_p2g_bwd(si, a_next, a_cur, svd);
    }
    // n >= 2 backward substeps (i = n-1 .. 0) with p2g.grad(i) and g2p.grad(i-1) fused; c = adjoint ping-pong parity at entry.
    // Inside a graph every frame i+1 was produced by the forward graph, so the fused kernel takes clamp masks / gather sums from
    // the stored frames; `next_ok` covers the leading unfused g2p.grad of substep n-1.
    // overlap (stream capture only): the grid pre-stage (restore + grid operator) of substep i-1 is captured on a forked branch
    // and runs beside the particle kernel and grid adjoint of substep i; the two grid sets alternate by substep parity:
    //   Pre(j) -> K(j) = [p2g.grad(j+1) +] g2p.grad(j) -> A(j) = grid adjoint(j) -> K(j-1);   Pre(j-2) waits for A(j) (same set).
    void enqueue_bwd_fused(int n, bool restore, bool next_ok, bool svd, int c, SlotRef (*mk)(const Engine*, int, int), bool overlap) {
        const bool direct = restore && slab_direct();          // (direct halo: the scatter of substep j and its grid adjoint use the parity-j buffer)
        struct PdlScope { bool& f; PdlScope(bool& f_) : f(f_) { f = true; } ~PdlScope() { f = false; } } pdl_scope(pdl_bwd);
        set_pdl_inhibit();
        if (!overlap) {
            enqueue_bwd_grid_pre(mk(this, 0, n - 1), mk(this, 2, n - 1), restore, sets[0], stream);
            select_gout(n - 1, direct);
            launch_g2p_bwd(mk(this, 0, n - 1), adj[c], adj[c ^ 1], sets[0], next_ok, tile_bwd, mk(this, 1, n - 1));
            enqueue_bwd_grid_adj(mk(this, 2, n - 1), sets[0]);
            for (int i = n - 1; i >= 1; i--) {
                enqueue_bwd_grid_pre(mk(this, 0, i - 1), mk(this, 2, i - 1), restore, sets[0], stream);
                select_gout(i - 1, direct);
                launch_bwd_fused(mk(this, 0, i), mk(this, 0, i - 1), adj[c], adj[c ^ 1], sets[0], svd);
                c ^= 1;
                enqueue_bwd_grid_adj(mk(this, 2, i - 1), sets[0]);
            }
            select_gout(0, false);
            launch_p2g_bwd(mk(this, 0, 0), adj[c], adj[c ^ 1], svd, tile_bwd);
            return;
        }
        cap_ev_used = 0;
        cudaEvent_t ev_fork = next_event();
        cudaEventRecord(ev_fork, stream);
        cudaStreamWaitEvent(side_stream, ev_fork, 0);
        // Pre(n-1) on the main stream, Pre(n-2) on the branch (the other set, free at graph entry)
        enqueue_bwd_grid_pre(mk(this, 0, n - 1), mk(this, 2, n - 1), true, sets[(n - 1) & 1], stream);
        cudaEvent_t ev_pre = next_event();
        enqueue_bwd_grid_pre(mk(this, 0, n - 2), mk(this, 2, n - 2), true, sets[(n - 2) & 1], side_stream);
        cudaEventRecord(ev_pre, side_stream);
        select_gout(n - 1, direct);
        launch_g2p_bwd(mk(this, 0, n - 1), adj[c], adj[c ^ 1], sets[(n - 1) & 1], next_ok, tile_bwd, mk(this, 1, n - 1));
        enqueue_bwd_grid_adj(mk(this, 2, n - 1), sets[(n - 1) & 1]);
        cudaEvent_t ev_adj = next_event();              // A(i) done, for the i of the coming iteration
        cudaEventRecord(ev_adj, stream);
        for (int i = n - 1; i >= 1; i--) {
            cudaStreamWaitEvent(stream, ev_pre, 0);     // Pre(i-1) done
            if (i - 2 >= 0) {                           // Pre(i-2) re-uses the set of substep i: wait for A(i)
                cudaStreamWaitEvent(side_stream, ev_adj, 0);
                enqueue_bwd_grid_pre(mk(this, 0, i - 2), mk(this, 2, i - 2), true, sets[(i - 2) & 1], side_stream);
                ev_pre = next_event();
                cudaEventRecord(ev_pre, side_stream);
            }
            select_gout(i - 1, direct);
            launch_bwd_fused(mk(this, 0, i), mk(this, 0, i - 1), adj[c], adj[c ^ 1], sets[(i - 1) & 1], svd);
            c ^= 1;
            enqueue_bwd_grid_adj(mk(this, 2, i - 1), sets[(i - 1) & 1]);
            ev_adj = next_event();
            cudaEventRecord(ev_adj, stream);
        }
        select_gout(0, false);
        launch_p2g_bwd(mk(this, 0, 0), adj[c], adj[c ^ 1], svd, tile_bwd);
    }
    // push my listed zone blocks of `grid` into the neighbours' inboxes, publish, wait for theirs
    void halo_exchange(const Vec4<T>* grid) {
        for (int side = 0; side < 2; side++)
            if (slab.has[side])
                k_halo_push<T><<<sparse_ctas(), kBlock, 0, stream>>>(cfg.n_grid, grid, d_list, d_nactive, slab.peer[side], slab.geom[side], slab.seq);
        k_halo_signal<<<1, 1, 0, stream>>>(slab.has[0] ? slab.peer[0] : nullptr, slab.has[1] ? slab.peer[1] : nullptr, slab.seq);
        k_halo_wait<<<1, 1, 0, stream>>>(slab.has[0] ? slab.inbox[0] : nullptr, slab.has[1] ? slab.inbox[1] : nullptr, slab.seq, slab.err);
    }
    void halo_add_inbox(Vec4<T>* grid) {
        for (int side = 0; side < 2; side++)
            if (slab.has[side])
                k_halo_add_inbox<T><<<sparse_ctas(), kBlock, 0, stream>>>(cfg.n_grid, grid, slab.inbox[side], slab.geom[side], d_list, d_nactive, slab.seq);
    }
    int own_lo() const { return slab.on ? slab.own_lo : 0; }
    int own_hi() const { return slab.on ? slab.own_hi : cfg.n_grid; }
    void launch_p2g(SlotRef si, SlotRef so, int store_F, bool mark = true) {
        unsigned char* fl = (sparse && mark) ? d_flags : nullptr;
        if (tile_scatter) {
            const int nbc = blocks(cfg.n_particles, cta);
            const size_t sm = tile_smem_bytes(cta);
            k_p2g_warp<T><<<nbc, cta, sm, stream>>>(P, frames, n_pad, si, so, store_F, material(), grid_in, fl, flush_mode, svd_store);
        } else {
            k_p2g<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, si, so, store_F, material(), grid_in, fl, svd_store);
        }
    }
    // (which: 0 slot_in base, 1 slot_out base, 2 pose frame base; rel) -> cursor-relative reference
    static SlotRef mk_cursor(const Engine* e, int which, int rel) { SlotRef r; r.cur = e->d_cursor; r.idx = which; r.rel = rel; return r; }
    static SlotRef abs_ref(int v) { SlotRef r; r.cur = nullptr; r.idx = 0; r.rel = v; return r; }
    SlotRef cur_ref(int idx, int rel) const { SlotRef r; r.cur = d_cursor; r.idx = idx; r.rel = rel; return r; }

    int substep_fwd(int si, int so, int pf) override {
        if (int r = check_slot(si)) return r;
        if (int r = check_slot(so)) return r;
        if (int r = check_pf(pf, 1)) return r;
        PLB_REQUIRE(si != so, "in-place substep");
        enqueue_fwd(abs_ref(si), abs_ref(so), abs_ref(pf));
        slot_written(so);
        stored[si] = store.vals != nullptr;
        fwd_ok[si] = (so == si + 1);
        svd_ok[si] = svd_store != nullptr;
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    int substep_bwd(int si, int pf) override {
        if (int r = check_slot(si)) return r;
        if (int r = check_pf(pf, 1)) return r;
        enqueue_bwd(abs_ref(si), abs_ref(pf), stored[si] && store.vals, fwd_ok[si] != 0, svd_ok[si] && svd_store, adj[cur], adj[cur ^ 1]);
        cur ^= 1;
        PLB_CUDA(cudaGetLastError());
        // (a forward pass made of plb_step_fwd calls may be differentiated substep by substep, like the reference's substep_grad loop:
        //  the adjoint of a re-sorted frame goes back to the previous ordering here as well)
        if (resort_q.count(si)) return unsort_frame(si);
        return PLB_OK;
    }

    // ---- whole env steps: one captured CUDA graph per (direction, n, adjoint parity, stored), replayed with a new cursor
    // the kernel sequence of one env step (cursor-relative frame references)
    void enqueue_step(const GraphKey& key) {
        const bool fused = fuse && tile_scatter && sparse && key.n >= 2;
        const bool restore = (key.stored & 1) != 0, next_ok = (key.stored & 2) != 0, svd = (key.stored & 4) != 0;
        if (key.dir == 0) {
            if (fused) enqueue_fwd_fused(key.n, &Engine::mk_cursor, env_list_mode());
            else for (int i = 0; i < key.n; i++) enqueue_fwd(cur_ref(0, i), cur_ref(1, i), cur_ref(2, i));
        } else {
            int c = key.parity;
            if (fused) enqueue_bwd_fused(key.n, restore, next_ok, svd, c, &Engine::mk_cursor, bwd_overlap && restore && (!slab.on || slab_fused()));
            else for (int i = key.n - 1; i >= 0; i--) { enqueue_bwd(cur_ref(0, i), cur_ref(2, i), restore, next_ok, svd, adj[c], adj[c ^ 1]); c ^= 1; }
        }
    }
    int launch_graph(const GraphKey& key, int slot0, int pf0) {
        if (prof_on) {
            // profiling (bench.py's live per-kernel times): the SAME kernel sequence as the graph, launched one by one with a
            // CUDA-event pair around every kernel, on the streams the graph's branches were captured from
            k_set_cursor<<<1, 1, 0, stream>>>(d_cursor, slot0, slot0 + 1, pf0);
            launches++;
            enqueue_step(key);
            cudaEvent_t join = next_event();            // (the overlapped backward leaves work on the side stream)
            cudaEventRecord(join, side_stream);
            cudaStreamWaitEvent(stream, join, 0);
            PLB_CUDA(cudaGetLastError());
            return PLB_OK;
        }
        auto it = graphs.find(key);
        if (it == graphs.end()) {
            cudaGraph_t g = nullptr;
            long long l0 = launches;
            PLB_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            enqueue_step(key);
            cudaError_t ce = cudaStreamEndCapture(stream, &g);
            graph_nodes[key] = launches - l0;
            launches = l0;
            if (ce != cudaSuccess) { err = std::string("graph capture: ") + cudaGetErrorString(ce); return PLB_ERR_CUDA; }
            cudaGraphExec_t ge = nullptr;
            PLB_CUDA(cudaGraphInstantiate(&ge, g, 0));
            cudaGraphDestroy(g);
            it = graphs.emplace(key, ge).first;
        }
        k_set_cursor<<<1, 1, 0, stream>>>(d_cursor, slot0, slot0 + 1, pf0);
        PLB_CUDA(cudaGraphLaunch(it->second, stream));
        // kernels + memset nodes inside the replayed graph (the capture counted them once into graph_nodes[key])
        launches += graph_nodes[key] + 1;
        return PLB_OK;
    }
    int step_fwd(int slot0, int pf0, int n) override {
        if (n <= 0) return PLB_OK;
        if (int r = check_slot(slot0)) return r;
        if (int r = check_slot(slot0 + n)) return r;
        if (int r = check_pf(pf0, n)) return r;
        if (!use_graphs || !sparse) {
            for (int i = 0; i < n; i++) if (int r = substep_fwd(slot0 + i, slot0 + i + 1, pf0 + i)) return r;
            return PLB_OK;
        }
        if (resort_applies(slot0)) { if (int r = resort_frame(slot0)) return r; }
        GraphKey key{0, n, 0, (store.vals != nullptr ? 1 : 0) | (env_list_mode() ? 4 : 0)};
        if (int r = launch_graph(key, slot0, pf0)) return r;
        if (!slot_order.empty()) for (int i = 1; i <= n; i++) slot_order[slot0 + i] = cur_order;
        for (int i = 0; i < n; i++) { stored[slot0 + i] = store.vals != nullptr; fwd_ok[slot0 + i] = 1; svd_ok[slot0 + i] = svd_store != nullptr; }
        stored[slot0 + n] = 0; fwd_ok[slot0 + n] = 0; svd_ok[slot0 + n] = 0;
        return PLB_OK;
    }
    int step_bwd(int slot0, int pf0, int n) override {
        if (n <= 0) return PLB_OK;
        if (int r = check_slot(slot0 + n - 1)) return r;
        if (int r = check_pf(pf0, n)) return r;
        int n_stored = 0, n_ok = 0, n_svd = 0;
        for (int i = 0; i < n; i++) { n_stored += stored[slot0 + i] ? 1 : 0; n_ok += fwd_ok[slot0 + i] ? 1 : 0; n_svd += svd_ok[slot0 + i] ? 1 : 0; }
        // the graphs take clamp masks / gather sums from the successor frames: every substep must have been run forward
        bool uniform = (n_stored == 0 || n_stored == n) && n_ok == n;
        if (!use_graphs || !sparse || !uniform) {
            for (int i = n - 1; i >= 0; i--) if (int r = substep_bwd(slot0 + i, pf0 + i)) return r;
            return PLB_OK;
        }
        GraphKey key{1, n, cur, ((n_stored == n && store.vals) ? 1 : 0) | 2 | ((n_svd == n && svd_store) ? 4 : 0)};
        if (int r = launch_graph(key, slot0, pf0)) return r;
        cur ^= (n & 1);
        if (resort_q.count(slot0)) return unsort_frame(slot0);
        return PLB_OK;
    }

    // ---------------------------------------------------------------- slab decomposition (multi-GPU)
    // The host (Python + torch.distributed over NCCL) moves zone planes between neighbours between the phases:
    //   fwd:  slab_fwd_p2g -> exchange grid_in zones -> slab_fwd_finish
    //   bwd:  slab_bwd_begin -> exchange g_out zones -> slab_bwd_finish
    //   loss: slab_loss_begin -> exchange grid_mass zones -> slab_loss_reduce -> all-reduce acc -> slab_loss_finish
    int slab_configure(int lo, int hi, int w, int has_left, int has_right) override {
        PLB_REQUIRE(sparse && tile_scatter, "slab mode needs kernel_variant 0");
        PLB_REQUIRE(lo % 4 == 0 && hi % 4 == 0 && w % 4 == 0 && w >= 4 && lo >= 0 && hi <= cfg.n_grid && lo < hi, "slab planes must be multiples of 4");
        PLB_REQUIRE((!has_left || lo - w >= 0) && (!has_right || hi + w <= cfg.n_grid), "halo outside the grid");
        PLB_REQUIRE(hi - lo >= 2 * w || !(has_left && has_right), "slab thinner than two halos");
        slab.on = true; slab.own_lo = lo; slab.own_hi = hi; slab.w = w;
        slab.has[0] = has_left != 0; slab.has[1] = has_right != 0;
        slab.zlo[0] = lo - w; slab.zhi[0] = lo + w; slab.zlo[1] = hi - w; slab.zhi[1] = hi + w;
        size_t plane = (size_t)cfg.n_grid * cfg.n_grid;
        for (int side = 0; side < 2; side++) {
            if (!slab.has[side]) continue;
            for (int which = 0; which < 3; which++) {
                if (slab.recv[which][side]) continue;
                size_t bytes = 2 * (size_t)w * plane * (which == 2 ? sizeof(T) : sizeof(Vec4<T>));
                PLB_CUDA(cudaMalloc(&slab.recv[which][side], bytes));
            }
        }
        if (slab.direct && !slab.g_out2) {
            const size_t gb = (size_t)n_nodes * sizeof(Vec4<T>);
            if (cudaMalloc(&slab.g_out2, gb) == cudaSuccess) cudaMemset(slab.g_out2, 0, gb);
            else { cudaGetLastError(); slab.g_out2 = nullptr; }
        }
        use_graphs = false;                    // phases are host-driven
        return PLB_OK;
    }
    // which: 0 grid_in, 1 g_out, 2 grid_mass; side: 0 left, 1 right; dir: 0 send (inside the grid), 1 recv (staging)
    int slab_buffer(int which, int side, int dir, void** ptr, long long* bytes) override {
        PLB_REQUIRE(slab.on && which >= 0 && which < 3 && side >= 0 && side < 2 && slab.has[side], "no such slab buffer");
        size_t plane = (size_t)cfg.n_grid * cfg.n_grid;
        size_t esz = which == 2 ? sizeof(T) : sizeof(Vec4<T>);
        *bytes = (long long)(2 * (size_t)slab.w * plane * esz);
        if (dir == 1) { *ptr = slab.recv[which][side]; return PLB_OK; }
        char* base = which == 0 ? (char*)grid_in : (which == 1 ? (char*)g_out : (char*)grid_mass);
        *ptr = base + (size_t)slab.zlo[side] * plane * esz;
        return PLB_OK;
    }
    int slab_fwd_p2g(int si, int so) override {
        if (int r = check_slot(si)) return r;
        if (int r = check_slot(so)) return r;
        PLB_REQUIRE(slab.on && si != so, "slab mode not configured");
        prof_begin(K_P2G);
        launch_p2g(abs_ref(si), abs_ref(so), 1);
        prof_end();
        launches++;
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    void halo_add(Vec4<T>* grid, int which) {
        for (int side = 0; side < 2; side++)
            if (slab.has[side]) {
                k_halo_add_listed<T><<<sparse_ctas(), kBlock, 0, stream>>>(cfg.n_grid, grid, (const Vec4<T>*)slab.recv[which][side], slab.zlo[side], slab.zhi[side], d_list, d_nactive);
                launches++;
            }
    }
    int slab_fwd_finish(int si, int so, int pf) override {
        if (int r = check_pf(pf, 1)) return r;
        PLB_REQUIRE(slab.on && store.vals, "slab mode needs the forward-grid store (call plb_sort_particles first)");
        prof_begin(K_GRID_FWD);
        for (int side = 0; side < 2; side++)
            if (slab.has[side]) {
                k_halo_mark<T><<<148, 256, 0, stream>>>(cfg.n_grid, (const Vec4<T>*)slab.recv[0][side], slab.zlo[side], slab.zhi[side], slab.own_lo, slab.own_hi, d_flags);
                launches++;
            }
        compact_blocks();
        halo_add(grid_in, 0);
        k_grid_fwd_sparse<T><<<sparse_ctas(), kBlock, 0, stream>>>(P, prims, d_traj, abs_ref(pf), grid_in, grid_out, 1, d_list, d_nactive, store, abs_ref(si), halo_in(0));
        prof_end(); prof_begin(K_G2P);
        k_g2p<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, abs_ref(si), abs_ref(so), grid_out);
        prof_end();
        launches += 2;
        slot_written(so);
        stored[si] = 1; fwd_ok[si] = (so == si + 1);
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    int slab_bwd_begin(int si, int pf) override {
        if (int r = check_slot(si)) return r;
        if (int r = check_pf(pf, 1)) return r;
        PLB_REQUIRE(slab.on && stored[si] && store.vals, "slab backward needs the stored forward grid of this slot");
        GridStore<T> nostore{nullptr, nullptr, nullptr, nullptr, 0};
        prof_begin(K_P2G_RECOMPUTE);
        k_restore_blocks<T><<<sparse_ctas(), kBlock, 0, stream>>>(cfg.n_grid, grid_in, d_list, d_nactive, store, abs_ref(si));
        prof_end(); prof_begin(K_GRID_FWD_RECOMPUTE);
        k_grid_fwd_sparse<T><<<sparse_ctas(), kBlock, 0, stream>>>(P, prims, d_traj, abs_ref(pf), grid_in, grid_out, 0, d_list, d_nactive, nostore, abs_ref(si), halo_in(0));
        prof_end();
        launch_g2p_bwd(abs_ref(si), adj[cur], adj[cur ^ 1], sets[0], fwd_ok[si] != 0);
        launches += 2;
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    int slab_bwd_finish(int si, int pf) override {
        PLB_REQUIRE(slab.on, "slab mode not configured");
        prof_begin(K_GRID_BWD);
        halo_add(g_out, 1);
        k_grid_bwd_sparse<T><<<sparse_ctas(), kBlock, 0, stream>>>(P, prims, d_traj, abs_ref(pf), grid_in, g_out, g_in, 1, d_prim_grad, d_list, d_nactive, own_lo(), own_hi());
        prof_end(); prof_begin(K_P2G_BWD);
        k_p2g_bwd<T, false><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, abs_ref(si), adj[cur], adj[cur ^ 1], material(), g_in, (T*)nullptr);
        prof_end();
        launches += 2;
        cur ^= 1;
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    int slab_loss_begin(int slot) override {
        if (int r = check_slot(slot)) return r;
        PLB_REQUIRE(slab.on && has_target, "slab loss needs slab mode and a target");
        k_loss_init<<<1, 32, 0, stream>>>(d_acc, lw.soft);
        PLB_CUDA(cudaMemsetAsync(grid_mass, 0, n_nodes * sizeof(T), stream));
        k_loss_mass_tile<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, slot, grid_mass);
        launches += 2;
        return PLB_OK;
    }
    int slab_loss_reduce(int slot, int pf) override {
        if (int r = check_pf(pf)) return r;
        size_t plane = (size_t)cfg.n_grid * cfg.n_grid;
        for (int side = 0; side < 2; side++)
            if (slab.has[side]) {
                long long cnt = 2LL * slab.w * plane;
                k_add_scalar<T><<<blocks(cnt), kBlock, 0, stream>>>(cnt, grid_mass + (size_t)slab.zlo[side] * plane, (const T*)slab.recv[2][side]);
                launches++;
            }
        long long own_nodes = (long long)(slab.own_hi - slab.own_lo) * plane;
        size_t off = (size_t)slab.own_lo * plane;
        int rb = (int)std::min<long long>((own_nodes + 255) / 256, 148 * 8);
        k_loss_reduce<T><<<rb, 256, 0, stream>>>(grid_mass + off, target + off, target_sdf + off, own_nodes, d_acc);
        k_loss_contact<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, prims, d_traj, pf, frames, n_pad, slot, d_acc, lw.soft);
        launches += 2;
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    // after the host all-reduced d_acc (sum [0..3], max [4], min [8..]) across ranks
    int slab_loss_finish(int slot, int pf, int backward, double* out8) override {
        if (!backward) {
            k_loss_finalize<T><<<1, 1, 0, stream>>>(prims, cfg.n_primitives, lw, d_acc, target_max, target_sum, d_acc + kAccN, d_acc + kAccN + 1);
            launches++;
            if (out8) {
                PLB_CUDA(cudaMemcpyAsync(out8, d_acc + kAccN + 1, 8 * sizeof(double), cudaMemcpyDeviceToHost, stream));
                PLB_CUDA(cudaStreamSynchronize(stream));
            }
        } else {
            k_loss_bwd<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, prims, d_traj, pf, frames, n_pad, slot, adj[cur], grid_mass, target,
                                                                          target_sdf, lw, d_acc, contact_all, d_prim_grad);
            launches++;
        }
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    // ---- peer-memory halo set-up: inbox allocation + CUDA IPC handles (exchanged by the host through torch.distributed)
    int alloc_inboxes() {
        if (slab.seq) return PLB_OK;
        const int nbx = cfg.n_grid / 4;
        for (int side = 0; side < 2; side++) {
            HaloGeom& g = slab.geom[side];
            g.zone_lo = slab.zlo[side]; g.zone_hi = slab.zhi[side];
            g.nzb = (2 * slab.w / 4) * nbx * nbx;
            g.stamps_off = 256;
            g.data_off = (256 + 2LL * g.nzb * (long long)sizeof(int) + 255) / 256 * 256;
            g.flags_off = g.data_off + 2LL * g.nzb * kBlkNodes * (long long)sizeof(Vec4<T>);
            slab.inbox_bytes = (size_t)g.flags_off + 2ULL * g.nzb;
            if (!slab.has[side]) continue;
            PLB_CUDA(cudaMalloc(&slab.inbox[side], slab.inbox_bytes));
            PLB_CUDA(cudaMemset(slab.inbox[side], 0, 256));
            PLB_CUDA(cudaMemset(slab.inbox[side] + g.stamps_off, 0xFF, 2ULL * g.nzb * sizeof(int)));
        }
        PLB_CUDA(cudaMalloc(&slab.seq, sizeof(int)));
        PLB_CUDA(cudaMemset(slab.seq, 0, sizeof(int)));
        PLB_CUDA(cudaMalloc(&slab.err, sizeof(int)));
        PLB_CUDA(cudaMemset(slab.err, 0, sizeof(int)));
        PLB_CUDA(cudaMalloc(&slab.done, sizeof(unsigned)));
        PLB_CUDA(cudaMemset(slab.done, 0, sizeof(unsigned)));
        PLB_CUDA(cudaMalloc(&slab.listed_stamp, n_blocks * sizeof(int)));
        PLB_CUDA(cudaMemset(slab.listed_stamp, 0xFF, n_blocks * sizeof(int)));
        return PLB_OK;
    }
    int slab_ipc_export(int side, void* handle64) override {
        PLB_REQUIRE(slab.on && side >= 0 && side < 2 && slab.has[side], "no neighbour on that side");
        if (int r = alloc_inboxes()) return r;
        cudaIpcMemHandle_t h;
        PLB_CUDA(cudaIpcGetMemHandle(&h, slab.inbox[side]));
        slab.exported = true;
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(handle64, &h, 64);
        return PLB_OK;
    }
    // side: the neighbour this handle came from (0 = my left neighbour's RIGHT inbox, 1 = my right neighbour's LEFT inbox)
    int slab_ipc_import(int side, const void* handle64) override {
        PLB_REQUIRE(slab.on && side >= 0 && side < 2 && slab.has[side], "no neighbour on that side");
        if (int r = alloc_inboxes()) return r;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, 64);
        void* ptr = nullptr;
        PLB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        slab.peer[side] = (char*)ptr;
        bool ready = true;
        for (int s2 = 0; s2 < 2; s2++) if (slab.has[s2] && !slab.peer[s2]) ready = false;
        if (ready) {
            slab.peer_ready = true;
            use_graphs = cfg.kernel_variant == 0 && !getenv("PLB_NO_GRAPHS");
            for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
            graphs.clear();
        }
        return PLB_OK;
    }
    // direct halo: which = 0/1 grid_in of even/odd substeps, 2/3 adjoint of grid_out of even/odd substeps
    int slab_ipc_export_grid(int which, void* handle64) override {
        PLB_REQUIRE(slab.on && which >= 0 && which < 4, "no such grid");
        PLB_REQUIRE(local_grid(which) != nullptr, "direct halo buffers are not allocated (PLB_SLAB_DIRECT=0 or PLB_BWD_OVERLAP=0)");
        cudaIpcMemHandle_t h;
        PLB_CUDA(cudaIpcGetMemHandle(&h, local_grid(which)));
        std::memcpy(handle64, &h, 64);
        return PLB_OK;
    }
    int slab_ipc_import_grid(int side, int which, const void* handle64) override {
        PLB_REQUIRE(slab.on && side >= 0 && side < 2 && slab.has[side] && which >= 0 && which < 4, "no neighbour on that side");
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, 64);
        void* ptr = nullptr;
        PLB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        slab.peer_grid[side][which] = (Vec4<T>*)ptr;
        drop_graphs();
        return PLB_OK;
    }
    int slab_ipc_close() override {
        PLB_CUDA(cudaStreamSynchronize(stream));
        drop_graphs();                         // (they hold the peer pointers by value)
        for (int side = 0; side < 2; side++)
            if (slab.peer[side]) { cudaIpcCloseMemHandle(slab.peer[side]); slab.peer[side] = nullptr; }
        for (int side = 0; side < 2; side++)
            for (int w = 0; w < 4; w++)
                if (slab.peer_grid[side][w]) { cudaIpcCloseMemHandle(slab.peer_grid[side][w]); slab.peer_grid[side][w] = nullptr; }
        slab.peer_ready = false; slab.ipc_closed = true;
        use_graphs = false;
        return PLB_OK;
    }
    // which: 0 loss accumulators (kAccN doubles), 1 primitive pose gradients (max_prim_frames*8*8 doubles)
    int device_buffer(int which, void** ptr, long long* bytes) override {
        if (which == 0) { *ptr = d_acc; *bytes = kAccN * (long long)sizeof(double); return PLB_OK; }
        if (which == 1) { *ptr = d_prim_grad; *bytes = (long long)(traj.size() * sizeof(double)); return PLB_OK; }
        err = "unknown device buffer"; return PLB_ERR_INVALID;
    }

    // ---------------------------------------------------------------- adjoint bookkeeping
    int zero_grads() override {
        PLB_CUDA(cudaMemsetAsync(adj[0], 0, (size_t)24 * n_pad * sizeof(T), stream));
        PLB_CUDA(cudaMemsetAsync(adj[1], 0, (size_t)24 * n_pad * sizeof(T), stream));
        PLB_CUDA(cudaMemsetAsync(d_prim_grad, 0, traj.size() * sizeof(double), stream));
        PLB_CUDA(cudaMemsetAsync(d_acc + kAccN, 0, sizeof(double), stream));
        scan_carry.reset();
        return PLB_OK;
    }
    int set_adjoint(const double* gx, const double* gv, const double* gF, const double* gC) override {
        double *dx, *dv, *dF, *dC;
        if (int r = upload_aos(gx, gv, gF, gC, &dx, &dv, &dF, &dC)) return r;
        k_pack_frame<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(cfg.n_particles, n_pad, adj[cur], d_perm, dx, dv, dF, dC);
        launches++;
        PLB_CUDA(cudaStreamSynchronize(stream));
        return PLB_OK;
    }
    int get_adjoint(double* gx, double* gv, double* gF, double* gC) override {
        if (int r = check_overflow()) return r;
        return download_aos(adj[cur], gx, gv, gF, gC);
    }
    int get_prim_grads(int pf0, int n, double* out) override {
        if (int r = check_pf(pf0, n - 1)) return r;
        std::vector<double> tmp((size_t)n * PLB_MAX_PRIM * 8);
        PLB_CUDA(cudaMemcpyAsync(tmp.data(), d_prim_grad + (size_t)pf0 * PLB_MAX_PRIM * 8, tmp.size() * sizeof(double),
                                 cudaMemcpyDeviceToHost, stream));
        PLB_CUDA(cudaStreamSynchronize(stream));
        for (int f = 0; f < n; f++)
            for (int k = 0; k < cfg.n_primitives; k++)
                std::memcpy(out + ((size_t)f * cfg.n_primitives + k) * 8, tmp.data() + ((size_t)f * PLB_MAX_PRIM + k) * 8, 8 * sizeof(double));
        return PLB_OK;
    }
    int check_overflow() {
        int ov = 0;
        PLB_CUDA(cudaMemcpyAsync(&ov, store.overflow, sizeof(int), cudaMemcpyDeviceToHost, stream));
        PLB_CUDA(cudaStreamSynchronize(stream));
        if (slab.err) {
            int he = 0;
            PLB_CUDA(cudaMemcpy(&he, slab.err, sizeof(int), cudaMemcpyDeviceToHost));
            if (he == 2) {
                cudaMemset(slab.err, 0, sizeof(int));
                err = "slab decomposition: a particle's stencil left this rank's owned planes + halo (particles do not migrate between slabs "
                      "inside an episode); results of this episode are invalid -- re-partition from the current state or use wider halos";
                return PLB_ERR_INVALID;
            }
            if (he) { cudaMemset(slab.err, 0, sizeof(int)); err = "slab halo: timed out waiting for a neighbour's push"; return PLB_ERR_CUDA; }
        }
        if (ov == 3) {
            PLB_CUDA(cudaMemsetAsync(store.overflow, 0, sizeof(int), stream));
            err = "env-step re-sort: the chunk table outgrew the launch grid of the captured graphs (the body spread over many more 4^3 blocks "
                  "than at plb_sort_particles); results of this episode are invalid -- run with PLB_RESORT=0";
            return PLB_ERR_NOMEM;
        }
        if (ov == 2) {
            PLB_CUDA(cudaMemsetAsync(store.overflow, 0, sizeof(int), stream));
            err = "PLB_ENV_LIST: the material left the block list of its env step (it moved more than one 4^3 block within the step); "
                  "results of this episode are invalid -- run without PLB_ENV_LIST";
            return PLB_ERR_NOMEM;
        }
        if (ov) {
            PLB_CUDA(cudaMemsetAsync(store.overflow, 0, sizeof(int), stream));
            err = "forward-grid store overflow: the material spread over more 4^3 blocks than reserved at the last "
                  "plb_sort_particles; gradients of this episode are invalid (re-sort the state or create the engine with kernel_variant=2)";
            return PLB_ERR_NOMEM;
        }
        return PLB_OK;
    }
    int get_action_grad(int n_steps, int S, double* out) override {
        int nf = n_steps * S;
        if (int r = check_pf(nf)) return r;
        if (int r = check_overflow()) return r;
        std::vector<double> g((size_t)(nf + 1) * PLB_MAX_PRIM * 8);
        PLB_CUDA(cudaMemcpyAsync(g.data(), d_prim_grad, g.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        PLB_CUDA(cudaStreamSynchronize(stream));
        std::fill(out, out + (size_t)n_steps * action_total, 0.0);
        kin::action_grad_scan(kdesc.data(), cfg.n_primitives, PLB_MAX_PRIM, traj.data(), vel.data(), g.data(), 0, 0, nf, S, action_off.data(),
                              action_total, out, 0);
        return PLB_OK;
    }

    // ---------------------------------------------------------------- policy path (plb/engine/nn/mlp.py, plb/optimizer/solver_nn.py)
    // With a state-feedback policy the action of env step t depends on the particle state and the poses at frame t*S, so its
    // gradient is needed while the backward sweep stands at that frame: the kinematics reverse scan is run one env step at a
    // time, carrying the adjoint that flows into the pose of frame t*S from everything after it (`scan_carry`), to which the
    // caller adds the policy's observation adjoint (plb_add_pose_adjoint) before the next (earlier) step.
    kin::ScanCarry scan_carry;
    int action_grad_step(int step, int S, double* out) override {
        const int f_lo = step * S, f_hi = (step + 1) * S;
        PLB_REQUIRE(step >= 0 && S > 0 && out != nullptr, "bad step");
        if (int r = check_pf(f_hi)) return r;
        if (int r = check_overflow()) return r;
        const size_t row = (size_t)PLB_MAX_PRIM * 8;
        std::vector<double> dev((size_t)(S + 1) * row), scratch((size_t)(S + 1) * row);
        PLB_CUDA(cudaMemcpyAsync(dev.data(), d_prim_grad + (size_t)f_lo * row, dev.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        PLB_CUDA(cudaStreamSynchronize(stream));
        const bool ok = kin::action_grad_step(kdesc.data(), cfg.n_primitives, traj.data(), vel.data(), dev.data(), step, S, action_off.data(),
                                              action_total, scan_carry, out, scratch.data());
        PLB_REQUIRE(ok, "plb_action_grad_step must walk the env steps in descending order");
        return PLB_OK;
    }
    int add_pose_adjoint(int k, const double* g8) override {
        PLB_REQUIRE(k >= 0 && k < cfg.n_primitives && g8 != nullptr, "primitive index");
        PLB_REQUIRE(scan_carry.frame >= 0, "plb_add_pose_adjoint needs a preceding plb_action_grad_step");
        for (int i = 0; i < 8; i++) scan_carry.v[(size_t)k * 8 + i] += g8[i];
        return PLB_OK;
    }
    // caller-order particle indices -> stored positions
    int* d_inv_perm = nullptr; bool inv_perm_valid = false;
    int* d_sel_idx = nullptr; double* d_sel_val = nullptr; int sel_cap = 0;
    int prepare_selection(const int* idx, int n) {
        PLB_REQUIRE(idx != nullptr && n > 0 && n <= cfg.n_particles, "bad particle selection");
        for (int i = 0; i < n; i++) PLB_REQUIRE(idx[i] >= 0 && idx[i] < cfg.n_particles, "particle index out of range");
        if (!d_inv_perm) PLB_CUDA(cudaMalloc(&d_inv_perm, n_pad * sizeof(int)));
        if (!inv_perm_valid) {
            k_invert_perm<<<blocks(cfg.n_particles), kBlock, 0, stream>>>(cfg.n_particles, d_perm, d_inv_perm);
            launches++;
            inv_perm_valid = true;
        }
        if (n > sel_cap) {
            cudaFree(d_sel_idx); cudaFree(d_sel_val);
            d_sel_idx = nullptr; d_sel_val = nullptr; sel_cap = 0;
            PLB_CUDA(cudaMalloc(&d_sel_idx, (size_t)n * sizeof(int)));
            PLB_CUDA(cudaMalloc(&d_sel_val, (size_t)n * 6 * sizeof(double)));
            sel_cap = n;
        }
        PLB_CUDA(cudaMemcpyAsync(d_sel_idx, idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, stream));
        return PLB_OK;
    }
    // x, v of the listed particles of frame `slot` (observation of the policy: mlp.py:68-76)
    int gather_particles(int slot, const int* idx, int n, double* x3, double* v3) override {
        if (int r = check_slot(slot)) return r;
        PLB_REQUIRE(x3 != nullptr && v3 != nullptr, "null output");
        if (int r = prepare_selection(idx, n)) return r;
        k_gather_xv<T><<<blocks(n), kBlock, 0, stream>>>(n, n_pad, frame_base(slot), d_sel_idx, d_inv_perm, d_sel_val);
        launches++;
        std::vector<double> h((size_t)n * 6);
        PLB_CUDA(cudaMemcpyAsync(h.data(), d_sel_val, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        PLB_CUDA(cudaStreamSynchronize(stream));
        for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) { x3[i * 3 + c] = h[(size_t)i * 6 + c]; v3[i * 3 + c] = h[(size_t)i * 6 + 3 + c]; }
        return PLB_OK;
    }
    // adds (gx, gv) to the CURRENT adjoint frame at the listed particles (distinct indices): adjoint of the observation
    int scatter_adjoint(const int* idx, int n, const double* gx3, const double* gv3) override {
        PLB_REQUIRE(gx3 != nullptr && gv3 != nullptr, "null input");
        if (int r = prepare_selection(idx, n)) return r;
        std::vector<double> h((size_t)n * 6);
        for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) { h[(size_t)i * 6 + c] = gx3[i * 3 + c]; h[(size_t)i * 6 + 3 + c] = gv3[i * 3 + c]; }
        PLB_CUDA(cudaMemcpyAsync(d_sel_val, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        k_scatter_adj_xv<T><<<blocks(n), kBlock, 0, stream>>>(n, n_pad, adj[cur], d_sel_idx, d_inv_perm, d_sel_val);
        launches++;
        PLB_CUDA(cudaStreamSynchronize(stream));      // h goes out of scope
        return PLB_OK;
    }

    // ---------------------------------------------------------------- loss
    int set_target(const double* density, const double* sdf) override {
        PLB_REQUIRE(density != nullptr, "density is NULL");
        double *d_den = nullptr, *d_sdf[2] = {nullptr, nullptr}, *d_near[2] = {nullptr, nullptr};
        int* d_changed = nullptr;
        int rc = PLB_OK;
        auto cleanup = [&]() { cudaFree(d_den); cudaFree(d_sdf[0]); cudaFree(d_sdf[1]); cudaFree(d_near[0]); cudaFree(d_near[1]); cudaFree(d_changed); };
        PLB_CUDA(cudaMalloc(&d_den, n_nodes * sizeof(double)));
        PLB_CUDA(cudaMemcpy(d_den, density, n_nodes * sizeof(double), cudaMemcpyHostToDevice));
        k_convert<T><<<blocks(n_nodes), kBlock, 0, stream>>>(n_nodes, d_den, target);
        launches++;
        target_max = 0; target_sum = 0;
        for (long long i = 0; i < n_nodes; i++) { target_max = std::max(target_max, density[i]); target_sum += density[i]; }
        if (cudaMalloc(&d_sdf[0], n_nodes * sizeof(double)) != cudaSuccess) { cleanup(); err = "cudaMalloc sdf"; return PLB_ERR_NOMEM; }
        if (sdf) {
            cudaMemcpy(d_sdf[0], sdf, n_nodes * sizeof(double), cudaMemcpyHostToDevice);
            k_convert<T><<<blocks(n_nodes), kBlock, 0, stream>>>(n_nodes, d_sdf[0], target_sdf);
            launches++;
        } else {
            if (cudaMalloc(&d_sdf[1], n_nodes * sizeof(double)) != cudaSuccess || cudaMalloc(&d_near[0], 3 * n_nodes * sizeof(double)) != cudaSuccess ||
                cudaMalloc(&d_near[1], 3 * n_nodes * sizeof(double)) != cudaSuccess || cudaMalloc(&d_changed, sizeof(int)) != cudaSuccess) {
                cleanup(); err = "cudaMalloc sdf sweep"; return PLB_ERR_NOMEM;
            }
            std::vector<double> inf((size_t)n_nodes, 1000.0);
            cudaMemcpy(d_sdf[0], inf.data(), n_nodes * sizeof(double), cudaMemcpyHostToDevice);
            cudaMemset(d_near[0], 0, 3 * n_nodes * sizeof(double));
            int c = 0;
            for (int it = 0; it < 2 * cfg.n_grid; it++) {
                cudaMemsetAsync(d_changed, 0, sizeof(int), stream);
                k_sdf_sweep<<<blocks(n_nodes, 128), 128, 0, stream>>>(cfg.n_grid, cfg.dx, d_den, d_sdf[c], d_near[c], d_sdf[c ^ 1], d_near[c ^ 1], d_changed);
                launches++;
                int changed = 1;
                cudaMemcpyAsync(&changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, stream);
                cudaStreamSynchronize(stream);
                c ^= 1;
                if (!changed) break;     // fixed point: further sweeps are identities
            }
            k_convert<T><<<blocks(n_nodes), kBlock, 0, stream>>>(n_nodes, d_sdf[c], target_sdf);
            launches++;
        }
        cudaError_t e = cudaStreamSynchronize(stream);
        cleanup();
        if (e != cudaSuccess) { err = cudaGetErrorString(e); return PLB_ERR_CUDA; }
        has_target = true;
        return rc;
    }
    int get_target_sdf(double* sdf) override {
        std::vector<T> tmp((size_t)n_nodes);
        PLB_CUDA(cudaMemcpy(tmp.data(), target_sdf, n_nodes * sizeof(T), cudaMemcpyDeviceToHost));
        for (long long i = 0; i < n_nodes; i++) sdf[i] = (double)tmp[i];
        return PLB_OK;
    }
    int set_loss_weights(double sdf, double density, double contact, int soft, int all) override {
        lw.sdf = sdf; lw.density = density; lw.contact = contact; lw.soft = soft; contact_all = all;
        return PLB_OK;
    }
    int loss_terms(int slot, int pf) {
        int nb = blocks(cfg.n_particles);
        k_loss_init<<<1, 32, 0, stream>>>(d_acc, lw.soft);
        PLB_CUDA(cudaMemsetAsync(grid_mass, 0, n_nodes * sizeof(T), stream));
        if (tile_scatter) k_loss_mass_tile<T><<<nb, kBlock, 0, stream>>>(P, frames, n_pad, slot, grid_mass);
        else k_loss_mass<T><<<nb, kBlock, 0, stream>>>(P, frames, n_pad, slot, grid_mass);
        int rb = (int)std::min<long long>((n_nodes + 255) / 256, 148 * 8);
        k_loss_reduce<T><<<rb, 256, 0, stream>>>(grid_mass, target, target_sdf, n_nodes, d_acc);
        k_loss_contact<T><<<nb, kBlock, 0, stream>>>(P, prims, d_traj, pf, frames, n_pad, slot, d_acc, lw.soft);
        launches += 4;
        return PLB_OK;
    }
    int loss_fwd(int slot, int pf, double* out8) override {
        if (int r = check_slot(slot)) return r;
        if (int r = check_pf(pf)) return r;
        PLB_REQUIRE(has_target, "no target density set");
        prof_begin(K_LOSS_FWD);
        if (int r = loss_terms(slot, pf)) return r;
        k_loss_finalize<T><<<1, 1, 0, stream>>>(prims, cfg.n_primitives, lw, d_acc, target_max, target_sum, d_acc + kAccN, d_acc + kAccN + 1);
        prof_end();
        launches++;
        PLB_CUDA(cudaGetLastError());
        if (out8) {
            PLB_CUDA(cudaMemcpyAsync(out8, d_acc + kAccN + 1, 8 * sizeof(double), cudaMemcpyDeviceToHost, stream));
            PLB_CUDA(cudaStreamSynchronize(stream));
        }
        return PLB_OK;
    }
    int loss_bwd(int slot, int pf) override {
        if (int r = check_slot(slot)) return r;
        if (int r = check_pf(pf)) return r;
        PLB_REQUIRE(has_target, "no target density set");
        prof_begin(K_LOSS_BWD);
        if (int r = loss_terms(slot, pf)) return r;
        k_loss_bwd<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, prims, d_traj, pf, frames, n_pad, slot, adj[cur], grid_mass, target,
                                                                      target_sdf, lw, d_acc, contact_all, d_prim_grad);
        prof_end();
        launches++;
        PLB_CUDA(cudaGetLastError());
        return PLB_OK;
    }
    int get_loss(double* v) override {
        PLB_CUDA(cudaMemcpyAsync(v, d_acc + kAccN, sizeof(double), cudaMemcpyDeviceToHost, stream));
        PLB_CUDA(cudaStreamSynchronize(stream));
        return PLB_OK;
    }
    int clear_loss() override { PLB_CUDA(cudaMemsetAsync(d_acc + kAccN, 0, sizeof(double), stream)); return PLB_OK; }

    // ---------------------------------------------------------------- debug
    int debug_get_grid(double* in4, double* out4) override {
        double* tmp = nullptr;
        PLB_CUDA(cudaMalloc(&tmp, n_nodes * 4 * sizeof(double)));
        const Vec4<T>* src[2] = {grid_in, grid_out};
        double* dst[2] = {in4, out4};
        for (int i = 0; i < 2; i++) {
            if (!dst[i]) continue;
            k_grid_to_double<T><<<blocks(n_nodes), kBlock, 0, stream>>>(n_nodes, src[i], tmp);
            launches++;
            cudaMemcpyAsync(dst[i], tmp, n_nodes * 4 * sizeof(double), cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
        }
        cudaFree(tmp);
        return PLB_OK;
    }
    int count_active(int slot, long long* n) override {
        if (int r = check_slot(slot)) return r;
        // scatter this frame's particles (no F store), count, then clear grid_in again
        k_p2g<T><<<blocks(cfg.n_particles), kBlock, 0, stream>>>(P, frames, n_pad, abs_ref(slot), abs_ref(slot), 0, material(), grid_in, nullptr, (T*)nullptr);
        PLB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream));
        k_count_active<T><<<blocks(n_nodes), kBlock, 0, stream>>>(n_nodes, grid_in, d_count);
        launches += 2;
        unsigned long long c = 0;
        PLB_CUDA(cudaMemcpyAsync(&c, d_count, sizeof(c), cudaMemcpyDeviceToHost, stream));
        PLB_CUDA(cudaMemsetAsync(grid_in, 0, n_nodes * sizeof(Vec4<T>), stream));
        PLB_CUDA(cudaStreamSynchronize(stream));
        *n = (long long)c;
        return PLB_OK;
    }
};

// ================================================================================================ C ABI
extern "C" {

int plb_abi_version(void) { return 1; }

int plb_create(const plb_config* cfg, const plb_primitive_desc* prims, plb_engine** out) {
    if (!cfg || !out) { g_create_error = "null argument"; return PLB_ERR_INVALID; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(ce) +
                         " (this engine has no CPU fallback)";
        return PLB_ERR_CUDA;
    }
    plb_engine* e = nullptr;
    if (cfg->dtype == PLB_F32) e = new Engine<float>();
    else if (cfg->dtype == PLB_F64) e = new Engine<double>();
    else { g_create_error = "dtype must be PLB_F32 or PLB_F64"; return PLB_ERR_INVALID; }
    int r = e->init(*cfg, prims);
    if (r != PLB_OK) { g_create_error = e->err; delete e; return r; }
    *out = e;
    return PLB_OK;
}
int plb_destroy(plb_engine* e) { delete e; return PLB_OK; }
const char* plb_last_error(const plb_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }
int plb_set_stream(plb_engine* e, void* s) { return e->set_stream(s); }
int plb_synchronize(plb_engine* e) { return e->synchronize(); }
int plb_set_materials(plb_engine* e, const double* mu, const double* lam, const double* ys) { return e->set_materials(mu, lam, ys); }
int plb_set_frame(plb_engine* e, int slot, const double* x, const double* v, const double* F, const double* C) { return e->set_frame(slot, x, v, F, C); }
int plb_get_frame(plb_engine* e, int slot, double* x, double* v, double* F, double* C) { return e->get_frame(slot, x, v, F, C); }
int plb_copy_frame(plb_engine* e, int s, int d) { return e->copy_frame(s, d); }
int plb_sort_particles(plb_engine* e, int slot) { return e->sort_particles(slot); }
int plb_frame_device_ptr(plb_engine* e, int slot, void** ptr, long long* n_pad, int* sb) { return e->frame_ptr(slot, ptr, n_pad, sb); }
int plb_set_primitive_state(plb_engine* e, int pf, int k, const double* s) { return e->set_prim_state(pf, k, s); }
int plb_get_primitive_state(plb_engine* e, int pf, int k, double* s) { return e->get_prim_state(pf, k, s); }
int plb_copy_primitive_frame(plb_engine* e, int s, int d) { return e->copy_prim_frame(s, d); }
int plb_set_softness(plb_engine* e, double s) { return e->set_softness(s); }
int plb_set_action(plb_engine* e, int step, int S, const double* a, int n) { return e->set_action(step, S, a, n); }
int plb_kinematics(plb_engine* e, int pf, int n) { return e->kinematics(pf, n); }
int plb_substep_fwd(plb_engine* e, int si, int so, int pf) { return e->substep_fwd(si, so, pf); }
int plb_step_fwd(plb_engine* e, int slot0, int pf0, int n) { return e->step_fwd(slot0, pf0, n); }
int plb_substep_bwd(plb_engine* e, int si, int pf) { return e->substep_bwd(si, pf); }
int plb_step_bwd(plb_engine* e, int slot0, int pf0, int n) { return e->step_bwd(slot0, pf0, n); }
int plb_zero_grads(plb_engine* e) { return e->zero_grads(); }
int plb_set_adjoint(plb_engine* e, const double* gx, const double* gv, const double* gF, const double* gC) { return e->set_adjoint(gx, gv, gF, gC); }
int plb_get_adjoint(plb_engine* e, double* gx, double* gv, double* gF, double* gC) { return e->get_adjoint(gx, gv, gF, gC); }
int plb_get_primitive_grads(plb_engine* e, int pf0, int n, double* out) { return e->get_prim_grads(pf0, n, out); }
int plb_get_action_grad(plb_engine* e, int n_steps, int S, double* out) { return e->get_action_grad(n_steps, S, out); }
int plb_action_grad_step(plb_engine* e, int step, int S, double* out) { return e->action_grad_step(step, S, out); }
int plb_add_pose_adjoint(plb_engine* e, int k, const double* g8) { return e->add_pose_adjoint(k, g8); }
int plb_gather_particles(plb_engine* e, int slot, const int* idx, int n, double* x3, double* v3) { return e->gather_particles(slot, idx, n, x3, v3); }
int plb_scatter_adjoint(plb_engine* e, const int* idx, int n, const double* gx3, const double* gv3) { return e->scatter_adjoint(idx, n, gx3, gv3); }
int plb_set_target(plb_engine* e, const double* d, const double* s) { return e->set_target(d, s); }
int plb_get_target_sdf(plb_engine* e, double* s) { return e->get_target_sdf(s); }
int plb_set_loss_weights(plb_engine* e, double s, double d, double c, int soft, int all) { return e->set_loss_weights(s, d, c, soft, all); }
int plb_loss_fwd(plb_engine* e, int slot, int pf, double* out8) { return e->loss_fwd(slot, pf, out8); }
int plb_loss_bwd(plb_engine* e, int slot, int pf) { return e->loss_bwd(slot, pf); }
int plb_get_loss(plb_engine* e, double* v) { return e->get_loss(v); }
int plb_clear_loss(plb_engine* e) { return e->clear_loss(); }
int plb_slab_configure(plb_engine* e, int lo, int hi, int w, int l, int r) { return e->slab_configure(lo, hi, w, l, r); }
int plb_slab_buffer(plb_engine* e, int which, int side, int dir, void** ptr, long long* bytes) { return e->slab_buffer(which, side, dir, ptr, bytes); }
int plb_slab_fwd_p2g(plb_engine* e, int si, int so) { return e->slab_fwd_p2g(si, so); }
int plb_slab_fwd_finish(plb_engine* e, int si, int so, int pf) { return e->slab_fwd_finish(si, so, pf); }
int plb_slab_bwd_begin(plb_engine* e, int si, int pf) { return e->slab_bwd_begin(si, pf); }
int plb_slab_bwd_finish(plb_engine* e, int si, int pf) { return e->slab_bwd_finish(si, pf); }
int plb_slab_loss_begin(plb_engine* e, int slot) { return e->slab_loss_begin(slot); }
int plb_slab_loss_reduce(plb_engine* e, int slot, int pf) { return e->slab_loss_reduce(slot, pf); }
int plb_slab_loss_finish(plb_engine* e, int slot, int pf, int backward, double* out8) { return e->slab_loss_finish(slot, pf, backward, out8); }
int plb_device_buffer(plb_engine* e, int which, void** ptr, long long* bytes) { return e->device_buffer(which, ptr, bytes); }
int plb_slab_ipc_export(plb_engine* e, int side, void* handle64) { return e->slab_ipc_export(side, handle64); }
int plb_slab_ipc_import(plb_engine* e, int side, const void* handle64) { return e->slab_ipc_import(side, handle64); }
int plb_slab_ipc_close(plb_engine* e) { return e->slab_ipc_close(); }
int plb_slab_ipc_export_grid(plb_engine* e, int which, void* handle64) { return e->slab_ipc_export_grid(which, handle64); }
int plb_slab_ipc_import_grid(plb_engine* e, int side, int which, const void* handle64) { return e->slab_ipc_import_grid(side, which, handle64); }
int plb_debug_get_grid(plb_engine* e, double* in4, double* out4) { return e->debug_get_grid(in4, out4); }
long long plb_launch_count(const plb_engine* e) { return e->launches; }
int plb_profile_enable(plb_engine* e, int on) {
    e->prof_collect();
    e->prof_on = on != 0;
    if (on) for (int i = 0; i < K_COUNT; i++) { e->prof_ms[i] = 0; e->prof_cnt[i] = 0; }
    return PLB_OK;
}
int plb_profile_read(plb_engine* e, int n, double* total_ms, long long* counts) {
    e->prof_collect();
    for (int i = 0; i < n && i < K_COUNT; i++) { total_ms[i] = e->prof_ms[i]; counts[i] = e->prof_cnt[i]; }
    return K_COUNT;
}
const char* plb_kernel_name(int kid) { return (kid >= 0 && kid < K_COUNT) ? kKernelNames[kid] : ""; }
int plb_count_active(plb_engine* e, int slot, long long* n) { return e->count_active(slot, n); }

}  // extern "C"
