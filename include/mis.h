/*
 * mis.h -- C-ABI of the B200-native meshless inflatable soft-body step.
 *
 * The reference (Megumi-X/meshless-inflatable-softbody) has no plugin / FFI
 * interface: its hot path is a set of module-level NVIDIA-Warp kernels launched
 * from a Python script.  This header is the boundary a replacement shared
 * library provides; every entry point cites the reference code it replaces.
 * The reference-side binding a maintainer would add is a ctypes stub, shown in
 * INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - `stream` is a cudaStream_t passed as void* (0 = the legacy default stream).
 *     Work is enqueued on it; no entry point synchronises the host unless its
 *     comment says so.
 *   - Pointers named *_dev are device pointers, *_host are host pointers (pinned
 *     for true asynchrony).  Per-particle arrays are in the CALLER's particle
 *     order (the order of x0 given to mis_create); the library keeps its own
 *     cell-sorted structure-of-arrays copies and un-permutes on export.
 *   - vec3 arrays are n*3 floats (AoS, as wp.array(dtype=vec3)); mat33 arrays are
 *     n*9 floats row-major (as wp.array(dtype=mat33)).
 *   - Return value: 0 = ok, negative = error (MIS_E_*); mis_last_error() gives a
 *     message for the calling thread.
 *   - One MisSim per stream, not re-entrant; distinct sims are independent.
 */
#ifndef MIS_H_
#define MIS_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MIS_OK              0
#define MIS_E_INVALID      -1   /* bad argument                                   */
#define MIS_E_CUDA         -2   /* a CUDA runtime call failed                     */
#define MIS_E_STATE        -3   /* call order violated (e.g. step before startup) */
#define MIS_E_UNSUPPORTED  -4   /* size/feature outside what this build supports  */

typedef struct MisSim MisSim;

/* Scene constants: the module-level constants of sim.py captured into its kernels. */
typedef struct MisParams {
    float h;                 /* sim.py:25   kernel radius, support 2h               */
    float damping;           /* sim.py:26                                           */
    float dt;                /* sim.py:65   time_step                               */
    float k_col;             /* sim.py:68   collision_penalty_stiffness             */
    float col_range;         /* sim.py:69   collision_range                         */
    float stiff_a, stiff_b;  /* sim.py:215  stress factor = a - b*ratio (200, 199)  */
    float tanh_k;            /* sim.py:110  ratio = 0.5*tanh(k*x)+0.5 (3)           */
    int   grid_x, grid_y, grid_z; /* sim.py:123-125 wp.HashGrid dims (cell_index export only) */
    /* sim_taichi.py deltas, 0 = sim.py behaviour */
    int   symmetric_pair;    /* sim_taichi.py:157  f_ij uses F_j                    */
    int   identity_rot;      /* sim_taichi.py:129  R = I                            */
    int   self_density;      /* sim_taichi.py:97   rho includes j == i              */
    int   euler;             /* sim_taichi.py:167-172 symplectic Euler              */
    int   no_contact;        /* sim_taichi.py has no ground penalty                 */
    /* tuning (0 = library default) */
    int   lanes_per_particle;/* 8, 16 or 32 lanes cooperate on one cluster          */
    int   keep_fields;       /* 1: also store A_pq each step (diagnostics export)   */
    int   graph_steps;       /* steps per captured CUDA graph chunk (0 = default 32, <0 = no graphs) */
    int   two_pass_deform;   /* 1: def_grad with the reference's two-loop order (sim.py:203-208)  */
    int   cluster_size;      /* 1, 2 or 4 consecutive cell-sorted particles share one union
                                neighbour list and every gathered record (0 = default 2)       */
    int   fp64;              /* 1: the scene's state and arithmetic are double (options.py:3 real = ti.f64): the
                                reference-order kernels of mis_ref.cuh step it; see mis_set_f64 / mis_get_f64 */
} MisParams;

typedef struct MisNeighborInfo {
    long long total_pairs;   /* sum of neighbour counts (directed pairs)            */
    int   max_neighbors;
    int   n;
    int   cell_min[3];       /* integer cell coordinate of the grid origin          */
    int   cell_dim[3];       /* dense cell table dims (x fastest)                   */
    float cell_width;        /* 2h                                                  */
    int   cluster_size;      /* particles per cluster (consecutive cell-sorted slots) */
    long long union_entries; /* sum of the clusters' union-list lengths             */
} MisNeighborInfo;

const char* mis_last_error(void);
/* library / build identification, e.g. "mis_b200 sm_100a <date>" */
const char* mis_version(void);

/* Replaces the allocation block sim.py:72-95 + wp.from_numpy(points_np) sim.py:85.
 * x0_dev: n*3 fp32 reference positions.  Copies x0; the caller keeps ownership.  */
int mis_create(int n, const float* x0_dev, const MisParams* params, void* stream, MisSim** out);
int mis_destroy(MisSim* sim);

/* Replaces wp.HashGrid(...).build(init_position, 2h), sim.py:123-127: cell binning of x0
 * (radix sort on Morton keys of the integer cell coordinates, dense cell table) plus the
 * static neighbour lists {j != i : |x0_i - x0_j|/h < 2} that the reference re-discovers
 * in every hash_grid_query (sim.py:161,178,203,224).  Synchronises the host (it sizes
 * the lists).  Idempotent: queries are centred on x0, so a rebuild yields identical
 * structures.  Called by mis_create; callable again to time the rebuild.            */
int mis_build_neighbors(MisSim* sim, void* stream);
int mis_get_neighbor_info(MisSim* sim, MisNeighborInfo* out);
/* Bit-exact checks against the reference structures.  Any pointer may be NULL.
 *   cell_index_dev[n]   : wp.HashGrid linear cell index of each particle (caller order)
 *   cell_coords_dev[3n] : integer cell coordinates int(p / cell_width) (caller order)
 *   perm_dev[n]         : caller id of the particle in sorted slot s.  Cells are contiguous slot
 *                         ranges; inside a cell particles follow a fine Morton curve, then caller id.
 *                         (hash_grid_point_id's order is the stable argsort of cell_index.)   */
int mis_export_cells(MisSim* sim, int* cell_index_dev, int* cell_coords_dev, int* perm_dev, void* stream);
/* dense cell table over cell_dim (x fastest): sorted-slot range [start, end) per cell */
int mis_export_cell_ranges(MisSim* sim, int* start_dev, int* end_dev, void* stream);
/* CSR neighbour lists in caller ids: offsets_dev[n+1] (row = caller id), nbr_dev[total_pairs]
 * (row entries in the library's walk order; compare as sets).                        */
int mis_export_neighbors(MisSim* sim, long long* offsets_dev, int* nbr_dev, void* stream);

/* --- control functions, sim.py:279-308.  Arrays are per particle, caller order. --- */
/* set_mass (sim.py:306-308): stores m and re-runs compute_v_i (sim.py:154-167).      */
int mis_set_mass(MisSim* sim, const float* mass_dev, void* stream);
/* set_youngs_modulus + set_poisson_ratio (sim.py:288-300): mu, lam from E, nu.       */
int mis_set_material(MisSim* sim, const float* youngs_dev, const float* poisson_dev, void* stream);
/* x -> ratio = 0.5*tanh(k x)+0.5, compute_ratio sim.py:107-110.                      */
int mis_set_design(MisSim* sim, const float* x_dev, void* stream);
/* external_forces array, sim.py:94,279-283 (n*3).                                    */
int mis_set_ext_force(MisSim* sim, const float* f_dev, void* stream);
int mis_set_ext_force_host(MisSim* sim, const float* f_host, void* stream);
/* free_points mask, sim.py:81,285-286 (n*3, component-wise multiplier).              */
int mis_set_dirichlet(MisSim* sim, const float* free_dev, void* stream);

/* startup kernel sim.py:261-266 (x = x0, v = v0) + the frame-0 force evaluation
 * sim.py:349-351.                                                                   */
int mis_startup(MisSim* sim, const float v0[3], void* stream);
/* resume from an arbitrary frame (x, v): re-primes the elastic force at x.           */
int mis_set_state(MisSim* sim, const float* x_dev, const float* v_dev, void* stream);
/* n_steps iterations of the loop body sim.py:352-358
 * (part_1 -> compute_A_pq -> compute_nabla_u -> compute_elastic_forces -> part_2).   */
int mis_step(MisSim* sim, int n_steps, void* stream);
/* position[f], velocity[f] of the current frame (sim.py:334,368-369), caller order.  */
int mis_get_state(MisSim* sim, float* x_dev, float* v_dev, void* stream);
int mis_get_state_host(MisSim* sim, float* x_host, float* v_host, void* stream);
/* Streaming export for per-frame consumers (visualize every 50th frame sim.py:393-395, targets sim.py:363-369): the un-permute
 * runs on `stream`, the device->host copies on a library-owned copy stream, so the following steps overlap the transfer.
 * Up to two exports may be in flight (double-buffered staging); x_host / v_host (pinned) are valid once
 * mis_wait_state_host(sim, pending_allowed) returns: it blocks until at most pending_allowed (0 or 1) exports are pending.
 * mis_set_ext_force_host uploads through the same copy stream.                                        */
int mis_get_state_host_async(MisSim* sim, float* x_host, float* v_host, void* stream);
int mis_wait_state_host(MisSim* sim, int pending_allowed);
/* Halo plumbing for slab-partitioned scenes (no reference counterpart: the reference is single-GPU).
 * part_1 of the NEXT step is fused into the force kernel, so after mis_step the positions of frame f+1
 * already exist; these two calls read / overwrite them for a subset of particles (caller ids, int32):
 * a rank sends the new positions of its boundary particles and overwrites its ghosts' before the next step. */
int mis_gather_next_positions(MisSim* sim, const int* ids_dev, int count, float* x_dev /* count*3 */, void* stream);
int mis_scatter_next_positions(MisSim* sim, const int* ids_dev, int count, const float* x_dev /* count*3 */, void* stream);
/* Ghost volumes of a slab-partitioned scene.  compute_v_i (sim.py:154-167) needs a particle's whole neighbourhood; an
 * outer ghost does not have it locally, so its owner's V (static: a function of x0 and m only) is written over the local
 * value once after mis_set_mass.  ids_dev: caller ids (int32), vol_dev[count].  Recomputes the static sums that use V.  */
int mis_set_volumes(MisSim* sim, const int* ids_dev, int count, const float* vol_dev, void* stream);
/* sorted slot (index into the library's cell-sorted arrays) of each caller id                   */
int mis_export_slots(MisSim* sim, const int* ids_dev, int count, int* slots_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused halo push over NVLink peer memory (no reference counterpart; replaces the gather / NCCL send-recv / scatter
 * round trip of mis_gather_next_positions + mis_scatter_next_positions).  The force kernel's part_1 epilogue
 * (sim.py:247-251) stores the new position of an owned boundary particle into its own array AND straight into the
 * ghost slot of every peer that mirrors it (P2P stores, overlapping the rest of the kernel); ghost slots are never
 * written locally.  One single-block kernel per step then publishes an epoch flag to each peer (release, system
 * scope) and waits for the peers' flags (acquire), so a whole chunk of steps -- pushes and synchronisation
 * included -- is one CUDA graph with no host or NCCL involvement.
 *   mis_halo_ipc_handles : 3 x 64 bytes = cudaIpcMemHandle_t of position buffer 0, position buffer 1, flag array.
 *   mis_halo_local_ptrs  : the same three as raw device pointers (peers living in the same process).
 *   mis_ipc_open / close : map a peer's handle into this process (cudaIpcOpenMemHandle, lazy peer access).
 *   mis_halo_connect     : peer_xv0/xv1[p] = peer p's position buffers mapped here; peer_flag[p] = address of MY entry
 *                          in peer p's flag array (its base + my index in its peer list; MIS_MAX_PEERS entries);
 *                          push_*: for each (owned caller id, peer index, slot on that peer) triple the particle is
 *                          mirrored there (at most two peers per particle); ghost_ids: local particles owned elsewhere;
 *                          ghost_layer (may be NULL = all 1): 1 = the ghost neighbours an owned particle (its R, S are recomputed
 *                          locally), 2 = outer layer (only its position is used).  Clusters made of ghosts skip the force gather,
 *                          clusters made of outer-layer ghosts also the deformation gather (their owner does that work).
 *                          All ranks must have finished their set-up (host barrier) before the first step after connect.
 *                          Epochs: every step and every re-prime (the first step after a setter / mis_set_state) publishes ONE
 *                          epoch and waits for the same epoch of each peer, so all ranks of a partition must issue the same
 *                          sequence of steps and state-changing calls (SPMD).  A rank that re-primes or steps alone runs ahead
 *                          of its peers' flags: its next wait then returns on stale ghosts or times out (mis_halo_status).
 *   mis_halo_status      : synchronises `stream`; err = 1 if a flag wait timed out (MIS_HALO_TIMEOUT_MS, default 20 s).  */
#define MIS_MAX_PEERS 4
int mis_halo_ipc_handles(MisSim* sim, unsigned char* out192);
int mis_halo_local_ptrs(MisSim* sim, void** xv0, void** xv1, void** flags);
int mis_ipc_open(const unsigned char* handle64, void** dev_ptr);
int mis_ipc_close(void* dev_ptr);
int mis_halo_connect(MisSim* sim, int n_peers, void* const* peer_xv0, void* const* peer_xv1, void* const* peer_flag,
                     int n_push, const int* push_ids_dev, const int* push_peer_dev, const int* push_slot_dev,
                     int n_ghost, const int* ghost_ids_dev, const int* ghost_layer_dev, void* stream);
int mis_halo_disconnect(MisSim* sim);
/* wait = 0: the per-step kernel only publishes this rank's epoch and does NOT wait for the peers'; the caller orders the ranks
 * itself (host synchronisation between steps).  For driving several ranks from one process on ONE GPU (tests), where kernels
 * that wait on one another must not be used: nothing guarantees that they run at the same time.  Default 1.       */
int mis_halo_set_wait(MisSim* sim, int wait);
int mis_halo_status(MisSim* sim, void* stream, int* err, long long* exchanges);

/* Per-particle fields of the current frame, caller order; any pointer may be NULL.
 * A_pq (needs keep_fields), R = U V^T, def_grad, S = compute_sigma, elastic force,
 * rho, volume  (sim.py:154-235).                                                    */
int mis_get_fields(MisSim* sim, float* A_dev, float* R_dev, float* F_dev, float* S_dev,
                   float* fel_dev, float* rho_dev, float* vol_dev, void* stream);
/* Evaluate the elastic force at an arbitrary configuration x (n*3) without touching the
 * state: compute_A_pq -> compute_nabla_u -> compute_elastic_forces once.              */
int mis_eval_forces(MisSim* sim, const float* x_dev, float* fel_dev, void* stream);
/* compute_loss, sim.py:269-273, forward value only: loss_dev[0] (fp64, device) += sum_i |x_i - xt_i|^2 + time_step |v_i - vt_i|^2
 * at the current frame; targets are n*3 fp32 in caller order (the position_{i}.npy / velocity_{i}.npy of sim.py:118-119).
 * Deterministic (fixed summation order).  The reverse pass (wp.Tape, sim.py:346-372) is out of scope.                       */
int mis_accumulate_loss(MisSim* sim, const float* target_x_dev, const float* target_v_dev, double* loss_dev, void* stream);
/* Gather mode of the two per-step neighbour-gather kernels (compute_A_pq / compute_nabla_u / compute_elastic_forces,
 * sim.py:170-235; same results to the fp32 summation-order floor):
 *   0  both passes: register-tiled clusters of 2 particles streaming a union neighbour list from global memory
 *      (k_deform_c, k_force_c)
 *   1  both passes: one CTA per hash-grid cell, its 27-cell neighbourhood staged in shared memory by TMA bulk copies, exact
 *      per-particle lists of uint16 tile offsets (k_deform_t + k_deform_fin, k_force_t)
 *   2  deformation pass as in 1, force pass as in 0 (the measured best on B200: the force pass gathers 80 B per pair and is
 *      bound by shared-memory bandwidth in mode 1, by L1 in mode 0 where a record serves both particles of a cluster)
 * Modes 1 and 2 return MIS_E_UNSUPPORTED if a neighbourhood of the scene holds more particles than the largest tile (the sim
 * then stays in mode 0).  Default: the environment variable MIS_GATHER (0 / 1 / 2) if set, else 2 when the scene fits.      */
int mis_set_gather_mode(MisSim* sim, int mode, void* stream);
/* out[0] mode in effect, [1] active cells, [2] largest tile (particles), [3] largest cell, [4] / [5] tile capacity of the
 * deform / force instantiation, [6] list blocks, [7] entries per block                                                       */
int mis_get_gather_info(MisSim* sim, int out[8]);
/* fp64 scenes (MisParams.fp64 = 1; sim_taichi.py runs in ti.f64, options.py:3).  The float setters / getters above keep
 * working (values converted exactly / rounded); these two move doubles, n x dim in caller order:
 *   what  0 x0 (3)  1 mass  2 youngs_modulus  3 poisson_ratio  4 design x  5 external_forces (3)  6 free_points (3)
 *         7 position (3)  8 velocity (3)                                              -- settable and gettable
 *         9 elastic_forces (3)  10 volume  11 rho  12 def_grad (9)  13 stress S (9)  14 rotation R (9)  15 A_pq (9)  -- get only
 * Not available for fp64 scenes: obstacle contact, halo plumbing, host streaming, mis_profile_step, mis_build_neighbors.      */
int mis_set_f64(MisSim* sim, int what, const double* src_dev, void* stream);
/* The scene constants of an fp64 scene in double (MisParams carries floats: 0.1f is not 0.1): c = { h, damping, time_step,
 * collision_penalty_stiffness, collision_range, stiffness_a, stiffness_b, tanh_k }.  The neighbour structure keeps the fp32 h
 * (a pair the two disagree on lies at q = 2, where W and nabla_W vanish).                                                    */
int mis_set_constants_f64(MisSim* sim, const double c[8]);
/* startup (sim.py:261-266 / sim_taichi.py:203-207) with the initial velocity in double                                       */
int mis_startup_f64(MisSim* sim, const double v0[3], void* stream);
int mis_get_f64(MisSim* sim, int what, double* dst_dev, void* stream);
/* Reverse pass of the rollout: diff_sim(compute_grad=True), sim.py:341-372 -- startup (with the initial velocity of the last
 * mis_startup), `frames` steps of the velocity-Verlet loop, compute_loss (sim.py:269-273) against target t (t = 0 ..
 * n_targets-1; n x 3 fp32 each, caller order, target t at offset t * 3n) at frame (frames / n_targets) * (t + 1), then the
 * adjoint sweep (wp.Tape.backward, sim.py:371).  Returns the loss (host) and d loss / d design x (sim.py:372 x.grad; n values,
 * caller order) as fp32 and/or fp64 (either pointer may be NULL).  The reference keeps every frame for its tape; here the
 * trajectory is checkpointed every `checkpoint_every` frames (0 = sqrt(frames)) and recomputed segment by segment.  Arithmetic:
 * the scene's precision (fp32, or fp64 with MisParams.fp64).  Synchronises the host; the scene must be re-started afterwards.   */
int mis_rollout_grad(MisSim* sim, int frames, int n_targets, const float* target_x_dev, const float* target_v_dev,
                     int checkpoint_every, double* loss_host, float* grad_dev, double* grad64_dev, void* stream);
/* number of kernels this sim has launched so far (bench.py's gpu_launches)           */
long long mis_launch_count(MisSim* sim);
/* Measurement aid: n_steps steps with a CUDA event pair around every kernel launch on
 * `stream`; returns the summed device milliseconds of the deform (compute_A_pq +
 * compute_nabla_u) and force (compute_elastic_forces + part_2/part_1) kernels.
 * Advances the state like mis_step.  Synchronises the host.                          */
int mis_profile_step(MisSim* sim, int n_steps, void* stream, double* ms_deform, double* ms_force);

/* ------------------------------------------------------------------------------------------
 * DeepSDF MLP (deepsdf.py:9-41, class DeepSDFWithCode; evaluated at sim.py:100).
 * n_layers Linear layers with dims[0] = 3, dims[1..n_layers-1] = H (network_size, a multiple of
 * 256), dims[n_layers] = 1; ReLU between layers; every Linear weight-normalised.  The arrays are
 * the tensors of the reference state dict, per Linear layer l (Sequential index 3 l):
 *   g_dev[l]    = network.{3l}.parametrizations.weight.original0   [out, 1]
 *   v_dev[l]    = network.{3l}.parametrizations.weight.original1   [out, in]   (W = g v / |v| per row)
 *   bias_dev[l] = network.{3l}.bias                                [out]
 * g_dev / v_dev / bias_dev are HOST arrays of n_layers DEVICE pointers.  Weights are copied
 * (packed into tensor-core tiles); the caller keeps ownership.  Synchronises the host.        */
typedef struct MisSdf MisSdf;
int mis_sdf_create(int n_layers, const int* dims, const float* const* g_dev, const float* const* v_dev,
                   const float* const* bias_dev, void* stream, MisSdf** out);
int mis_sdf_destroy(MisSdf* sdf);
/* sdf(points): DeepSDFWithCode.forward (deepsdf.py:40-41).  points_dev is n*3 fp32; if xform_host
 * is non-NULL it holds 12 floats (A row-major 3x3, then t) and the network sees A (p - t), the
 * inverse of the asset placement p_world = p_model @ R + lift (sim.py:46-52); NULL = model-space
 * input.  sdf_dev[n] receives the values.  grad_dev (n*3, may be NULL) receives d sdf / d p in the
 * frame of `points` by forward differences of step fd_eps in model space (3 extra evaluations). */
int mis_sdf_query(MisSdf* sdf, const float* points_dev, int n, const float* xform_host,
                  float* sdf_dev, float* grad_dev, float fd_eps, void* stream);
/* Hidden-layer kernel choice (tests / tuning): 0 = automatic (few rows or a device-side row count -- the per-step contact query:
 * ALL hidden layers in one cooperative launch, split-K over 8-CTA clusters with a distributed-shared-memory reduction and a
 * device-wide barrier between layers; bulk queries: persistent 128 x 256 tiles, one launch per layer), 1 = split-K, one launch per
 * layer, 2 = always big tiles, 3 = always the one-launch chain.  All are correct for any row count.              */
int mis_sdf_set_gemm_path(MisSdf* sdf, int path);
/* kernels launched so far by this network / of which tcgen05 GEMM launches                    */
long long mis_sdf_launch_count(MisSdf* sdf, long long* gemm_launches);
/* Measurement aid: device milliseconds of `reps` back-to-back hidden-layer GEMMs (layer 1) on
 * m rows of scratch activations (CUDA events on `stream`).  Synchronises the host.            */
int mis_sdf_profile_gemm(MisSdf* sdf, int m, int reps, void* stream, double* ms_total);

/* Per-step obstacle contact (EXTENSION: the reference evaluates DeepSDF once at start-up,
 * sim.py:100, and its per-step contact is the ground plane, sim.py:238-244).  Generalises
 * compute_collision_penalty: delta = range - sdf(p_model), f = delta^2 * k_col * n,
 * n = grad sdf / |grad sdf| in world space, added next to the ground penalty in part_1 / part_2.
 * xform_host as in mis_sdf_query; bbox_host = model-space min xyz, max xyz of the region where
 * sdf may be < range (broad phase: particles outside are skipped).  sdf = NULL disables.       */
int mis_set_sdf_contact(MisSim* sim, MisSdf* sdf, const float* xform_host, const float* bbox_host,
                        float fd_eps, void* stream);
/* obstacle-contact force at the current frame's positions (n*3, caller order)                  */
int mis_get_contact_force(MisSim* sim, float* f_dev, void* stream);
/* most recent step: count[0] = particles that passed the broad phase (MLP evaluated),
 * count[1] = particles inside the contact band (normal evaluated, force applied).  Synchronises. */
int mis_get_contact_count(MisSim* sim, void* stream, int count[2]);

#ifdef __cplusplus
}
#endif
#endif /* MIS_H_ */
