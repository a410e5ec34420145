/* plb_b200.h -- C ABI of the B200-native differentiable MPM engine.
 *
 * The reference (hzaskywalker/PlasticineLab) has no FFI: its engine is Python + Taichi JIT kernels.  This
 * header declares the boundary a replacement engine exports underneath the reference's Python object surface
 * (SURVEY.md section 8b).  Each entry point names the reference interface it replaces (file:line relative to
 * the reference checkout).  Everything is `extern "C"`, plain pointers and sizes, no torch/CUDA types:
 * streams travel as `void*` (a cudaStream_t), device buffers are owned by the engine.
 *
 * Conventions
 *   - every function returns 0 on success, a negative plb_status on failure; plb_last_error() gives the text;
 *   - host arrays are float64, row-major, laid out exactly like the numpy arrays of the reference
 *     (x,v: [N][3]; F,C: [N][3][3]; primitive state: 7 or 8 doubles = position, quaternion wxyz[, gap]);
 *   - `slot` indexes the engine's particle-frame storage (0 .. max_frames-1); `pf` indexes primitive frames
 *     (0 .. max_prim_frames-1).  The reference uses one index f for both (fields shaped [max_steps, ...],
 *     plb/engine/mpm_simulator.py:35-38, plb/engine/primitive/primive_base.py:36-37); keeping them apart lets
 *     the host run env-step checkpointing (plb/optimizer/long_term_gradient.ipynb cell 4) in a small window;
 *   - calls are asynchronous on the engine's stream unless they return data to the host;
 *   - one host thread per engine; one engine per GPU.
 */
#ifndef PLB_B200_H
#define PLB_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plb_engine plb_engine;

typedef enum {
    PLB_OK = 0,
    PLB_ERR_INVALID = -1,      /* bad argument / state */
    PLB_ERR_CUDA = -2,         /* CUDA runtime error */
    PLB_ERR_NOMEM = -3,
    PLB_ERR_UNSUPPORTED = -4
} plb_status;

enum { PLB_F32 = 0, PLB_F64 = 1 };

/* primitive shapes: plb/engine/primitive/primitives.py:17-257 */
enum { PLB_SPHERE = 0, PLB_CAPSULE = 1, PLB_ROLLINGPIN = 2, PLB_CHOPSTICKS = 3, PLB_CYLINDER = 4,
       PLB_TORUS = 5, PLB_BOX = 6 };

#define PLB_MAX_PRIMITIVES 8
#define PLB_MAX_ACTION_DIM 7

/* One rigid manipulator; fields = the reference's per-primitive cfg (primive_base.py:208-224, primitives.py). */
typedef struct {
    int    type;                 /* PLB_SPHERE ... */
    double params[4];            /* sphere: radius | capsule/rollingpin/chopsticks: h, r | cylinder: h, r |
                                    torus: tx, ty | box: size xyz */
    double friction;
    double init_state[8];        /* init_pos(3), init_rot(4), init_gap */
    double lower_bound[3];
    double upper_bound[3];
    int    action_dim;           /* 0 = static */
    double action_scale[PLB_MAX_ACTION_DIM];
    double minimal_gap;          /* chopsticks */
} plb_primitive_desc;

/* Simulator configuration = derived constants of MPMSimulator.__init__ (mpm_simulator.py:6-50). */
typedef struct {
    int    dtype;                /* PLB_F32 (production) or PLB_F64 (parity mode; the reference is f64 only) */
    int    n_particles;
    int    n_grid;               /* int(128 * quality / 2) */
    int    substeps;             /* int(2e-3 // dt), informational */
    int    max_frames;           /* particle frame slots to allocate */
    int    max_prim_frames;      /* primitive trajectory length */
    double dt, dx, p_vol, p_mass;
    double E, nu;                /* -> mu, lam (uniform material) */
    double yield_stress;
    double ground_friction;
    double gravity[3];
    int    n_primitives;
    int    device;               /* CUDA device ordinal */
    int    kernel_variant;       /* 0 = default (fastest available), 1 = simple reference kernels */
} plb_config;

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
/* replaces TaichiEnv.__init__/initialize + ti.init (plb/engine/taichi_env.py:6,10-58) */
int  plb_create(const plb_config* cfg, const plb_primitive_desc* prims, plb_engine** out);
int  plb_destroy(plb_engine* e);
const char* plb_last_error(const plb_engine* e);      /* e may be NULL: error of the last failed plb_create */
int  plb_set_stream(plb_engine* e, void* cuda_stream);
int  plb_synchronize(plb_engine* e);
int  plb_abi_version(void);

/* ---- particle state ------------------------------------------------------------------------------------- */
/* MPMSimulator.initialize material fill, mpm_simulator.py:53-57; NULL = keep the uniform value */
int  plb_set_materials(plb_engine* e, const double* mu, const double* lam, const double* yield_stress);
/* MPMSimulator.setframe / readframe / get_x / get_v, mpm_simulator.py:282-363 (any pointer may be NULL) */
int  plb_set_frame(plb_engine* e, int slot, const double* x, const double* v, const double* F, const double* C);
int  plb_get_frame(plb_engine* e, int slot, double* x, double* v, double* F, double* C);
/* MPMSimulator.copyframe, mpm_simulator.py:303-312 (particles only; primitives: plb_copy_primitive_frame) */
int  plb_copy_frame(plb_engine* e, int src_slot, int dst_slot);
/* Re-orders the particles stored in `slot` spatially (by 4^3 grid block, then cell) so that consecutive particles
   share stencil nodes; every other slot becomes stale.  Host-visible particle order is unchanged (the engine keeps the
   permutation and applies it in plb_set_frame / plb_get_frame / plb_set_materials / plb_*_adjoint).  No reference
   counterpart: the reference never reorders particles (plb/engine/mpm_simulator.py:157-184 scatters in sampling order). */
int  plb_sort_particles(plb_engine* e, int slot);
/* device-side access for torch interop: pointer to the frame's first scalar, padded particle count, scalar size */
int  plb_frame_device_ptr(plb_engine* e, int slot, void** ptr, long long* n_pad, int* scalar_bytes);

/* ---- primitives ----------------------------------------------------------------------------------------- */
/* Primitive.set_state / get_state, primive_base.py:143-150 (+ Chopsticks gap, primitives.py:133-146) */
int  plb_set_primitive_state(plb_engine* e, int pf, int k, const double* state8);
int  plb_get_primitive_state(plb_engine* e, int pf, int k, double* state8);
int  plb_copy_primitive_frame(plb_engine* e, int src_pf, int dst_pf);
/* Primitives.set_softness, primitives.py:303-305 */
int  plb_set_softness(plb_engine* e, double softness);
/* Primitives.set_action -> set_velocity, primitives.py:289-293, primive_base.py:184-198: clips to [-1,1], fills the
   per-substep velocities of frames [step*S, (step+1)*S) */
int  plb_set_action(plb_engine* e, int step, int n_substeps, const double* action, int action_len);
/* forward_kinematics for frames pf .. pf+n-1 -> pf+1 .. pf+n (primive_base.py:117-121, primitives.py:66-80,94-98),
   then uploads the poses */
int  plb_kinematics(plb_engine* e, int pf, int n);

/* ---- simulation ----------------------------------------------------------------------------------------- */
/* MPMSimulator.substep(s), mpm_simulator.py:245-257: state[slot_in] -> state[slot_out] with poses pf, pf+1 */
int  plb_substep_fwd(plb_engine* e, int slot_in, int slot_out, int pf);
/* n consecutive substeps slot0+i -> slot0+i+1, poses pf0+i (MPMSimulator.step's loop, mpm_simulator.py:373-374): one CUDA graph
   launch.  If frame slot0 was produced by an earlier plb_step_fwd, its particles are first re-sorted spatially in place (like
   plb_sort_particles, nothing read back); the engine keeps the permutation and plb_step_bwd / plb_substep_bwd of slot0 put the
   adjoint frame, frame slot0 and the per-particle materials back into the previous order, so callers never see it: every getter
   and setter maps through the ordering the frame is stored in.  PLB_RESORT=0 in the environment switches the re-sort off. */
int  plb_step_fwd(plb_engine* e, int slot0, int pf0, int n);
/* MPMSimulator.substep_grad(s), mpm_simulator.py:260-278: adjoint(frame s+1) -> adjoint(frame s); adds the pose
   adjoints of frames pf, pf+1 to the primitive-gradient buffer */
int  plb_substep_bwd(plb_engine* e, int slot_in, int pf);
int  plb_step_bwd(plb_engine* e, int slot0, int pf0, int n);      /* substeps slot0+n-1 .. slot0 */

/* ---- adjoint state (ti.Tape bookkeeping, plb/optimizer/solver.py:36) -------------------------------------- */
int  plb_zero_grads(plb_engine* e);                    /* particle adjoint, primitive gradients, loss value */
int  plb_set_adjoint(plb_engine* e, const double* gx, const double* gv, const double* gF, const double* gC);
int  plb_get_adjoint(plb_engine* e, double* gx, double* gv, double* gF, double* gC);
/* d loss / d pose for primitive frames [pf0, pf0+n): out[n][n_primitives][8] */
int  plb_get_primitive_grads(plb_engine* e, int pf0, int n, double* out);
/* Primitives.get_grad(n) (primitives.py:295-301): chains the pose adjoints through forward_kinematics.grad and
   set_velocity.grad; out[n_steps][sum action_dim] */
int  plb_get_action_grad(plb_engine* e, int n_steps, int n_substeps, double* out);

/* ---- policy path: state-feedback policies differentiated through the simulator ------------------------------
   (plb/engine/nn/mlp.py:68-134 reads x, v of every (N // 200)-th particle and the primitive poses at frame t*S,
   writes action_buffer[t]; plb/optimizer/solver_nn.py:33-43 replays it under ti.Tape.)
   plb_gather_particles : x[n][3], v[n][3] of the listed particles (caller-order indices) of frame `slot`.
   plb_scatter_adjoint  : adds gx[n][3], gv[n][3] to the CURRENT adjoint frame at the listed (distinct) particles --
                          the adjoint of that observation, once the backward sweep stands at the observed frame.
   plb_action_grad_step : like plb_get_action_grad for ONE env step (out[sum action_dim]); must be called for the env
                          steps in descending order, each after plb_step_bwd of that step; keeps the pose adjoint that
                          flows into frame step*S from everything after it.
   plb_add_pose_adjoint : adds g8 = d loss / d (position(3), rotation(4), gap) of primitive k at that frame (the
                          observation of the primitive state) to the carried adjoint. */
int  plb_gather_particles(plb_engine* e, int slot, const int* idx, int n, double* x3, double* v3);
int  plb_scatter_adjoint(plb_engine* e, const int* idx, int n, const double* gx3, const double* gv3);
int  plb_action_grad_step(plb_engine* e, int step, int n_substeps, double* out);
int  plb_add_pose_adjoint(plb_engine* e, int k, const double* g8);

/* ---- loss (plb/engine/losses/loss.py) --------------------------------------------------------------------- */
/* Loss.load_target_density + update_target (loss.py:46-66,81-106).  density: [n_grid^3] float64.
   sdf may be NULL: the engine then builds it on the device with the reference's sweep. */
int  plb_set_target(plb_engine* e, const double* density, const double* sdf);
int  plb_get_target_sdf(plb_engine* e, double* sdf);
/* Loss.set_weights (loss.py:68-72); contact_grad_all = 1 reproduces the reference's autodiff of atomic_min */
int  plb_set_loss_weights(plb_engine* e, double sdf, double density, double contact, int soft_contact,
                          int contact_grad_all);
/* Loss.compute_loss_kernel(f) + iou (loss.py:186-208,239-254).  Adds the step loss to the accumulated loss.
   out8 (may be NULL = stay asynchronous): accumulated loss, contact, density, sdf, iou, step loss, 0, 0 */
int  plb_loss_fwd(plb_engine* e, int slot, int pf, double* out8);
/* Loss.compute_loss_kernel.grad(f) (loss.py:210-237): seeds the particle adjoint (frame = slot) and pose grads */
int  plb_loss_bwd(plb_engine* e, int slot, int pf);
int  plb_get_loss(plb_engine* e, double* accumulated);     /* loss.loss[None] */
int  plb_clear_loss(plb_engine* e);                        /* Loss.clear_loss (loss.py:181-183) */

/* ---- multi-GPU slab decomposition (no reference counterpart: the reference is single-device) -------
 * One engine per GPU holds the particles whose base cell lies in its slab of grid planes [own_lo, own_hi) along axis 0
 * (planes are contiguous in memory).  Stencils reach into the neighbours' slabs, so the planes [b-w, b+w) around each
 * boundary b are computed by both neighbours as partial sums and exchanged by the host (NCCL send/recv of the buffers
 * returned by plb_slab_buffer) between the two halves of a substep:
 *   forward : plb_slab_fwd_p2g   -> exchange which=0 (grid_in zones)   -> plb_slab_fwd_finish
 *   backward: plb_slab_bwd_begin -> exchange which=1 (g_out zones)     -> plb_slab_bwd_finish
 *   loss    : plb_slab_loss_begin-> exchange which=2 (grid_mass zones) -> plb_slab_loss_reduce
 *             -> all-reduce of buffer 0 (sum [0..3], max [4], min [8..15]) -> plb_slab_loss_finish
 * The grid operator runs redundantly inside the zones; a plane's nodes contribute their pose gradients on the rank that
 * owns the plane; buffer 1 (pose gradients) is sum-all-reduced before plb_get_action_grad. */
int  plb_slab_configure(plb_engine* e, int own_lo, int own_hi, int halo_w, int has_left, int has_right);
int  plb_slab_buffer(plb_engine* e, int which, int side, int dir, void** ptr, long long* bytes);
int  plb_slab_fwd_p2g(plb_engine* e, int slot_in, int slot_out);
int  plb_slab_fwd_finish(plb_engine* e, int slot_in, int slot_out, int pf);
int  plb_slab_bwd_begin(plb_engine* e, int slot_in, int pf);
int  plb_slab_bwd_finish(plb_engine* e, int slot_in, int pf);
int  plb_slab_loss_begin(plb_engine* e, int slot);
int  plb_slab_loss_reduce(plb_engine* e, int slot, int pf);
int  plb_slab_loss_finish(plb_engine* e, int slot, int pf, int backward, double* out8);
int  plb_device_buffer(plb_engine* e, int which, void** ptr, long long* bytes);
/* Peer-memory halo (preferred): every rank exports a CUDA-IPC handle of its per-side inbox, the host hands it to the
 * neighbour on that side (torch.distributed all_gather_object), the neighbour imports it.  Once all neighbours are
 * imported, plb_substep_fwd/bwd and plb_step_fwd/bwd do the halo themselves: after the scatter each rank stores ONLY its
 * listed 4^3 blocks inside the zone into the neighbour's inbox (P2P stores over NVLink), stamps them with a sequence
 * number, publishes the number in a flag, waits for its own flag, adds what arrived -- all in-stream, so whole env steps
 * are CUDA graphs again.  Inside plb_step_fwd/bwd the grid kernels do all of this themselves (push, interior blocks, wait,
 * zone blocks; no extra launches).  handle64: 64 bytes (cudaIpcMemHandle_t). */
int  plb_slab_ipc_export(plb_engine* e, int side, void* handle64);
int  plb_slab_ipc_import(plb_engine* e, int side, const void* handle64);
/* Unmaps the neighbours' inboxes (cudaIpcCloseMemHandle) and returns to the host-driven halo.  Every rank must call this --
 * and the ranks must meet at a barrier -- before any rank destroys its engine: CUDA forbids freeing exported memory while an
 * importer still maps it. */
int  plb_slab_ipc_close(plb_engine* e);
/* Direct halo (on top of the inbox exchange above; optional): the ranks also map each other's scatter targets -- which = 0 / 1:
 * grid_in of even / odd substeps, 2 / 3: adjoint of grid_out of even / odd substeps -- and the scatter kernels add their zone
 * contributions straight into the neighbours' copies (red.global over NVLink) while they add them locally; an exchange is then
 * only the completion flag published by the last CTA of the scatter kernel, and the grid kernels wait for it.  Export all four,
 * import the four of each neighbour (side: 0 = left neighbour's, 1 = right neighbour's).  Without these calls the engine pushes
 * zone blocks into the inboxes instead. */
int  plb_slab_ipc_export_grid(plb_engine* e, int which, void* handle64);
int  plb_slab_ipc_import_grid(plb_engine* e, int side, int which, const void* handle64);

/* ---- introspection for tests / profiling ------------------------------------------------------------------ */
/* copies the dense grids of the last substep: any of in4/out4 may be NULL; [n_grid^3][4] float64 */
int  plb_debug_get_grid(plb_engine* e, double* in4, double* out4);
/* number of kernels this engine has launched since creation */
long long plb_launch_count(const plb_engine* e);
/* per-kernel device time, measured with CUDA events on the engine's stream around every launch while enabled.
   plb_profile_read: synchronises, fills total_ms[i] / counts[i] for kernel ids 0..n-1, returns the number of ids. */
int  plb_profile_enable(plb_engine* e, int on);
int  plb_profile_read(plb_engine* e, int n, double* total_ms, long long* counts);
const char* plb_kernel_name(int kernel_id);
/* nodes with mass > 1e-12 after the most recent forward P2G of plb_count_active (synchronous) */
int  plb_count_active(plb_engine* e, int slot, long long* n_active);

#ifdef __cplusplus
}
#endif
#endif /* PLB_B200_H */
