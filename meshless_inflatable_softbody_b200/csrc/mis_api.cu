// mis_api.cu -- the C-ABI of include/mis.h: owns the cell-sorted structure-of-arrays
// state of one scene and enqueues the kernels of mis_sort.cuh, mis_neighbors.cuh, mis_cluster.cuh and mis_sdf.cuh.
#include "../../include/mis.h"
#include "mis_math.cuh"
#include "mis_neighbors.cuh"
#include "mis_sort.cuh"
#include "mis_cluster.cuh"
#include "mis_tile.cuh"
#include "mis_tilebuild.cuh"
#include "mis_sdf_host.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

using namespace mis;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(MIS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)
#define CK_LAUNCH() CK(cudaGetLastError())

struct MisSim {
    int n = 0;
    MisParams p{};
    Consts c{};
    int G = 8, C = 2;                 // lanes per cluster, particles per cluster
    bool merge_lists = false;         // MIS_MERGE_LISTS=1: cluster lists by merging the members' exact lists (2.3x faster rebuild; the appended order costs the
                                      // force kernel 3-4 % in gather locality, so the 27-cell walk order stays the default)
    bool pair_cells = true;           // MIS_PAIR_CELLS=0: keep the plain in-cell Morton order (A/B measurement)
    // caller-order copies
    float* x0_orig = nullptr;
    int* coords = nullptr;            // Warp's truncated cell coordinates (exported)
    int* bcoords = nullptr;           // floor() coordinates: what the cell table and the Morton keys are built from
    int* cell_index = nullptr;
    // sort
    uint32_t* keys = nullptr;
    uint32_t* subkey = nullptr;
    int sub_bits = 0;
    uint32_t* perm = nullptr;
    int* inv_perm = nullptr;
    RadixSortTemp rs;
    int* bounds_dev = nullptr;
    // cell table
    int cell_min[3] = {0, 0, 0}, cell_dim[3] = {0, 0, 0};
    int ncells = 0, ncells_cap = 0;
    int* cell_start = nullptr;
    int* cell_end = nullptr;
    int* cell_lin_sorted = nullptr;
    // lists
    uint32_t* nbr_count = nullptr;
    unsigned long long* nbr_start = nullptr;
    unsigned long long* scan_tmp = nullptr;
    uint32_t* nbr = nullptr;
    long long nbr_cap = 0, total_pairs = 0;
    int* max_k_dev = nullptr;
    int max_k = 0;
    // cluster union lists (what the step kernels stream)
    uint32_t* cl_count = nullptr;
    unsigned long long* cl_start = nullptr;
    uint32_t* cl = nullptr;
    long long cl_cap = 0, cl_total = 0;
    float d2_limit = 0.f;
    // per-SM cluster queues of the persistent force kernel (k_force_p)
    int nsm = 0, force_bps = 0;
    int* sm_first = nullptr;
    int* sm_ctr = nullptr;
    bool force_persist = true;        // MIS_FORCE_PERSIST=0: always one block per 16 clusters (k_force_c), the round-1 launch shape
    // shared-memory cell tiles (mis_tile.cuh): gather mode 1
    struct Tile {
        bool want_d = false, want_f = false, ok = false;      // tiles for the deformation pass / the force pass
        uint32_t *wkey = nullptr, *order = nullptr;
        int n_active = 0, max_tile = 0, max_own = 0, cap_d = 0, cap_f = 0;
        int* tab = nullptr; long long tab_cap = 0;
        uint32_t* flag = nullptr; unsigned long long* pos = nullptr;
        uint32_t* nblocks = nullptr; unsigned long long* blk_start = nullptr; uint32_t* t_off = nullptr;
        unsigned short* lists = nullptr; long long lists_cap = 0, total_blocks = 0;
        float4* AB = nullptr;
        int* max_dev = nullptr;
        uint32_t* bits = nullptr; long long bits_cap = 0; int W = 0;       // neighbour bitmasks over the tile (mis_tilebuild.cuh)
        bool tab_ok = false;                                               // the tile table of the current binning exists
    } tile;
    // cell-sorted state
    float4 *x0m = nullptr, *xv[2] = {nullptr, nullptr}, *vel = nullptr, *f1 = nullptr, *fel = nullptr;
    float4 *fext = nullptr, *freem = nullptr, *matl = nullptr, *RS = nullptr, *Fd = nullptr, *Ks = nullptr;
    float* Apq = nullptr;
    float4* scratch4 = nullptr;       // eval / host staging
    double* loss_partial = nullptr;   // block partial sums of mis_accumulate_loss
    float* stage = nullptr;           // n*6 device staging for host-buffer variants
    int cur = 0;
    float* raw_in = nullptr;          // 3n: Young's modulus, Poisson ratio, design x as given (slot order): inputs of the precision-templated engine
    double v0[3] = {0.0, 0.0, 0.0};  // initial velocity of the last mis_startup / mis_startup_f64
    void* ref32 = nullptr;            // RefEngine<float>*  (mis_rollout_grad on an fp32 scene)
    void* ref64 = nullptr;            // RefEngine<double>* (MisParams.fp64: the whole scene steps in double)
    cudaError_t enqueue_err = cudaSuccess;    // first error of a launch inside an enqueue_* helper (reported by the calling entry point)
    bool built = false, mass_set = false, material_set = false, started = false, dirty = true;
    bool forces_only = false;         // dirty because of fext / free_points alone: the stored elastic force is still valid
    long long launches = 0;
    // CUDA graph cache: one executable per (steps in the chunk, ping-pong index at its start)
    struct StepGraph { cudaGraphExec_t exec = nullptr; int steps = 0, cur = -1; long long launches = 0; };
    std::vector<StepGraph> graphs;
    cudaStream_t graph_stream = nullptr;
    // DeepSDF obstacle contact (extension)
    MisSdf* sdf = nullptr;
    SdfXform sdf_xf{};
    float3 sdf_lo{}, sdf_hi{};
    float sdf_eps = 1e-3f;
    int *con_idx = nullptr, *con_idx2 = nullptr;
    int* con_count = nullptr;                 // [0] broad-phase candidates, [1] particles in the contact band, [2] FD rows, [3] overflow flag
    int con_cap = 0;                          // rows one pass of the chain can take (MIS_CONTACT_ROWS, default 32768, at most n)
    float *con_pts = nullptr, *con_pts2 = nullptr, *con_s0 = nullptr;
    float4* fcon = nullptr;
    // the contact chain only reads the positions of the new frame, like k_deform_c: it runs beside it on a forked,
    // higher-priority stream (a parallel branch of the step graph) and joins before k_force_c
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_chain = nullptr;
    bool serial_contact = false;
    bool contact_first = true;                // MIS_CONTACT_FIRST=0: launch the deformation kernel at once instead of behind the chain's first layer
    bool contact_split = false;               // the force gather stores fel (MODE_EVAL) and k_integrate follows the contact chain: the chain
                                              // overlaps BOTH gather kernels (MIS_CONTACT_SPLIT=0/1; default: scenes of <= 400 k particles)
    // fused halo push (slab-partitioned scenes)
    int2* push = nullptr;                       // per-slot destination codes
    unsigned* halo_mem = nullptr;               // [0..MIS_MAX_PEERS) flags written by the peers, [MIS_MAX_PEERS] epoch, [MIS_MAX_PEERS + 1] error
    int halo_peers = 0;
    bool halo_on = false;
    bool halo_wait = true;
    float4* peer_xv[2][MIS_MAX_PEERS] = {};
    unsigned* peer_flag[MIS_MAX_PEERS] = {};
    unsigned long long halo_timeout_ns = 20000000000ull;
    long long exchanges = 0;              // MIS_SERIAL_CONTACT=1: deform -> contact -> force in one stream (A/B measurement)
    // host <-> device streaming on a second stream (copy engines overlap the step kernels): double-buffered staging
    cudaStream_t copy_stream = nullptr;         // device -> host
    cudaStream_t up_stream = nullptr;           // host -> device (separate, so an upload never queues behind a download that waits for a step)
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_up_used[2] = {nullptr, nullptr}, ev_exp[2] = {nullptr, nullptr}, ev_down[2] = {nullptr, nullptr};
    float* fstage[2] = {nullptr, nullptr};      // n*3 each: external force uploads
    float* sstage[2] = {nullptr, nullptr};      // n*6 each: exported position + velocity
    long long up_i = 0, down_i = 0;
};

#include "mis_ref_host.cuh"
static inline RefEngine<double>* R64(MisSim* s) { return (RefEngine<double>*)s->ref64; }
static inline RefEngine<float>* R32(MisSim* s) { return (RefEngine<float>*)s->ref32; }
#define NO_FP64(what) do { if (s && s->ref64) return fail(MIS_E_UNSUPPORTED, what " is not available for an fp64 scene (MisParams.fp64)"); } while (0)

static int ensure_copy_stream(MisSim* s) {
    if (s->copy_stream) return MIS_OK;
    CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CK(cudaEventCreateWithFlags(&s->ev_up[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s->ev_up_used[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->ev_exp[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s->ev_down[k], cudaEventDisableTiming));
        CK(cudaMalloc((void**)&s->fstage[k], 3 * (size_t)s->n * sizeof(float) + 64));
        CK(cudaMalloc((void**)&s->sstage[k], 6 * (size_t)s->n * sizeof(float) + 64));
    }
    return MIS_OK;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T) + 64); }

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e && e[0] ? atoi(e) : dflt; }

extern "C" const char* mis_last_error(void) { return g_err.c_str(); }
extern "C" const char* mis_version(void) { return "mis_b200 sm_100a (" __DATE__ ")"; }

static float find_d2_limit(float h) {
    // smallest float d2 with !(sqrtf(d2)/h < 2): the predicate is monotone in d2
    auto pred = [h](float d2) { volatile float r = sqrtf(d2); volatile float q = r / h; return q < 2.f; };
    uint32_t lo = 0, hi;
    float hif = 16.f * h * h;
    memcpy(&hi, &hif, 4);
    while (hi - lo > 1) {          // pred(lo) true, pred(hi) false; positive floats order like their bits
        uint32_t mid = lo + (hi - lo) / 2;
        float mf; memcpy(&mf, &mid, 4);
        if (pred(mf)) lo = mid; else hi = mid;
    }
    float out; memcpy(&out, &hi, 4);
    return out;
}

static void make_consts(MisSim* s) {
    const MisParams& p = s->p;
    Consts& c = s->c;
    const float pi = 3.14159265358979323846f;
    c.h = p.h; c.inv_h = 1.f / p.h;
    c.sigma = 1.f / (pi * p.h * p.h * p.h);
    c.grad_c1 = c.sigma / (p.h * p.h);
    c.grad_c2 = 0.75f * c.sigma / p.h;
    c.sigma4 = 0.25f * c.sigma; c.grad_q1 = 3.f * c.sigma / p.h; c.grad_q2 = 0.75f * c.sigma / p.h;
    c.dt = p.dt; c.half_dt2 = 0.5f * p.dt * p.dt; c.damping = p.damping;
    c.k_col = p.k_col; c.col_range = p.col_range;
    c.stiff_a = p.stiff_a; c.stiff_b = p.stiff_b;
    c.identity_rot = p.identity_rot; c.euler = p.euler; c.no_contact = p.no_contact; c.symmetric_pair = p.symmetric_pair;
}

static View make_view(MisSim* s) {
    View v;
    v.n = s->n;
    v.x0m = s->x0m; v.xcur = s->xv[s->cur]; v.xnext = s->xv[s->cur ^ 1];
    v.vel = s->vel; v.f1 = s->f1; v.fel = s->fel; v.fext = s->fext; v.freem = s->freem; v.matl = s->matl;
    v.RS = s->RS; v.Fd = s->Fd; v.Ks = s->Ks; v.Apq = s->p.keep_fields ? s->Apq : nullptr;
    v.cl_start = s->cl_start; v.cl = s->cl;
    v.fcon = s->sdf ? s->fcon : nullptr;
    v.push = s->halo_on ? s->push : nullptr;
    for (int p = 0; p < 4; p++) v.peer_x[p] = s->halo_on ? s->peer_xv[s->cur ^ 1][p] : nullptr;
    return v;
}

static void drop_graph(MisSim* s) {
    for (auto& g : s->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    s->graphs.clear();
}

static int ref64_attach(MisSim* s, cudaStream_t st);
// ------------------------------------------------------------------ create / destroy
extern "C" int mis_create(int n, const float* x0_dev, const MisParams* params, void* stream, MisSim** out) {
    if (!out || !params || !x0_dev || n <= 0) return fail(MIS_E_INVALID, "mis_create: bad argument");
    if (!(params->h > 0.f) || !(params->dt > 0.f)) return fail(MIS_E_INVALID, "mis_create: h and dt must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    MisSim* s = new MisSim();
    s->n = n; s->p = *params;
    if (s->p.stiff_a == 0.f && s->p.stiff_b == 0.f) { s->p.stiff_a = 200.f; s->p.stiff_b = 199.f; }
    if (s->p.tanh_k == 0.f) s->p.tanh_k = 3.f;
    if (s->p.grid_x < 1) s->p.grid_x = 1;
    if (s->p.grid_y < 1) s->p.grid_y = 1;
    if (s->p.grid_z < 1) s->p.grid_z = 1;
    int G = s->p.lanes_per_particle ? s->p.lanes_per_particle : 8;
    if (G != 8 && G != 16 && G != 32) { delete s; return fail(MIS_E_INVALID, "lanes_per_particle must be 8, 16 or 32"); }
    s->G = G;
    int Cs = s->p.cluster_size ? s->p.cluster_size : 2;
    if (Cs != 1 && Cs != 2 && Cs != 4) { delete s; return fail(MIS_E_INVALID, "cluster_size must be 1, 2 or 4"); }
    s->C = Cs;
    { const char* e = getenv("MIS_PAIR_CELLS"); if (e && e[0] == '0') s->pair_cells = false; }
    { const char* e = getenv("MIS_MERGE_LISTS"); if (e && e[0] == '1') s->merge_lists = true; }
    { const char* e = getenv("MIS_FORCE_PERSIST"); if (e && e[0] == '0') s->force_persist = false; }
    {   // gather mode: 2 (default) = shared-memory tiles for the deformation pass, cluster kernel for the force pass
        const char* e = getenv("MIS_GATHER");
        const int m = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
        s->tile.want_d = m >= 1; s->tile.want_f = m == 1;
    }
    make_consts(s);
    s->d2_limit = find_d2_limit(s->p.h);
    const size_t N = (size_t)n;
#define ALLOC(ptr, cnt) do { cudaError_t e_ = dalloc(&(ptr), (cnt)); if (e_ != cudaSuccess) { int r_ = fail(MIS_E_CUDA, std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_)); mis_destroy(s); return r_; } } while (0)
    ALLOC(s->x0_orig, 3 * N); ALLOC(s->coords, 3 * N); ALLOC(s->bcoords, 3 * N); ALLOC(s->cell_index, N);
    ALLOC(s->keys, N); ALLOC(s->subkey, N); ALLOC(s->perm, N); ALLOC(s->inv_perm, N);
    s->rs.nblocks = nblk(n, RS_TILE);
    ALLOC(s->rs.keys_alt, N); ALLOC(s->rs.vals_alt, N);
    ALLOC(s->rs.hist, (size_t)RS_RADIX * s->rs.nblocks); ALLOC(s->rs.hist_scanned, (size_t)RS_RADIX * s->rs.nblocks + 1);
    ALLOC(s->rs.tile_tmp, (size_t)nblk((long long)RS_RADIX * s->rs.nblocks, SCAN_TILE) + 1);
    ALLOC(s->bounds_dev, 8); ALLOC(s->max_k_dev, 2);
    ALLOC(s->cell_lin_sorted, N);
    ALLOC(s->nbr_count, N); ALLOC(s->nbr_start, N + 1); ALLOC(s->scan_tmp, (size_t)nblk(n, SCAN_TILE) + 1);
    ALLOC(s->cl_count, N); ALLOC(s->cl_start, N + 1);
    ALLOC(s->x0m, N); ALLOC(s->xv[0], N); ALLOC(s->xv[1], N); ALLOC(s->vel, N); ALLOC(s->f1, N); ALLOC(s->fel, N);
    ALLOC(s->fext, N); ALLOC(s->freem, N); ALLOC(s->matl, N); ALLOC(s->RS, 4 * N); ALLOC(s->Fd, 3 * N); ALLOC(s->Ks, 3 * N);
    ALLOC(s->scratch4, 2 * N); ALLOC(s->stage, 6 * N); ALLOC(s->raw_in, 3 * N);
    if (s->p.keep_fields) ALLOC(s->Apq, 9 * N);
#undef ALLOC
    cudaMemsetAsync(s->x0m, 0, N * sizeof(float4), st);
    cudaMemsetAsync(s->xv[0], 0, N * sizeof(float4), st);
    cudaMemsetAsync(s->xv[1], 0, N * sizeof(float4), st);
    cudaMemsetAsync(s->vel, 0, N * sizeof(float4), st);
    cudaMemsetAsync(s->f1, 0, N * sizeof(float4), st);
    cudaMemsetAsync(s->fel, 0, N * sizeof(float4), st);
    cudaMemsetAsync(s->fext, 0, N * sizeof(float4), st);
    cudaMemsetAsync(s->matl, 0, N * sizeof(float4), st);
    cudaMemsetAsync(s->RS, 0, 4 * N * sizeof(float4), st);
    cudaMemsetAsync(s->Fd, 0, 3 * N * sizeof(float4), st);
    cudaMemsetAsync(s->Ks, 0, 3 * N * sizeof(float4), st);
    {   // free_points = 1 (sim.py:81)
        std::vector<float4> ones(N, make_float4(1.f, 1.f, 1.f, 0.f));
        cudaError_t e = cudaMemcpyAsync(s->freem, ones.data(), N * sizeof(float4), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { int r = fail(MIS_E_CUDA, std::string("init free_points: ") + cudaGetErrorString(e)); mis_destroy(s); return r; }
    }
    cudaError_t e = cudaMemcpyAsync(s->x0_orig, x0_dev, 3 * N * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { int r = fail(MIS_E_CUDA, std::string("copy x0: ") + cudaGetErrorString(e)); mis_destroy(s); return r; }
    int rc = mis_build_neighbors(s, stream);
    if (rc != MIS_OK) { mis_destroy(s); return rc; }
    if (s->p.fp64) { rc = ref64_attach(s, st); if (rc != MIS_OK) { mis_destroy(s); return rc; } }
    *out = s;
    return MIS_OK;
}

extern "C" int mis_destroy(MisSim* s) {
    if (!s) return MIS_OK;
    drop_graph(s);
    if (s->side_stream) {
        cudaStreamSynchronize(s->side_stream);
        cudaEventDestroy(s->ev_fork); cudaEventDestroy(s->ev_join); if (s->ev_chain) cudaEventDestroy(s->ev_chain);
        cudaStreamDestroy(s->side_stream);
    }
    if (s->copy_stream) {
        cudaStreamSynchronize(s->copy_stream);
        for (int k = 0; k < 2; k++) {
            cudaEventDestroy(s->ev_up[k]); cudaEventDestroy(s->ev_up_used[k]); cudaEventDestroy(s->ev_exp[k]); cudaEventDestroy(s->ev_down[k]);
            cudaFree(s->fstage[k]); cudaFree(s->sstage[k]);
        }
        cudaStreamDestroy(s->copy_stream);
        if (s->up_stream) { cudaStreamSynchronize(s->up_stream); cudaStreamDestroy(s->up_stream); }
    }
    void* ptrs[] = {s->loss_partial, s->push, s->halo_mem, s->con_idx, s->con_idx2, s->con_count, s->con_pts, s->con_pts2, s->con_s0, s->fcon, s->x0_orig, s->coords, s->bcoords, s->cell_index, s->keys, s->subkey, s->Ks, s->perm, s->inv_perm, s->rs.keys_alt, s->rs.vals_alt,
                    s->rs.hist, s->rs.hist_scanned, s->rs.tile_tmp, s->bounds_dev, s->max_k_dev, s->cell_start, s->cell_end,
                    s->cell_lin_sorted, s->nbr_count, s->nbr_start, s->scan_tmp, s->nbr, s->cl_count, s->cl_start, s->cl, s->x0m, s->xv[0], s->xv[1], s->vel,
                    s->f1, s->fel, s->fext, s->freem, s->matl, s->RS, s->Fd, s->Apq, s->scratch4, s->stage};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (s->ref32) { R32(s)->release(); delete R32(s); }
    if (s->ref64) { R64(s)->release(); delete R64(s); }
    if (s->raw_in) cudaFree(s->raw_in);
    if (s->sm_first) cudaFree(s->sm_first);
    if (s->sm_ctr) cudaFree(s->sm_ctr);
    void* tptrs[] = {s->tile.bits, s->tile.wkey, s->tile.order, s->tile.tab, s->tile.flag, s->tile.pos, s->tile.nblocks, s->tile.blk_start, s->tile.t_off, s->tile.lists, s->tile.AB, s->tile.max_dev};
    for (void* p : tptrs) if (p) cudaFree(p);
    delete s;
    return MIS_OK;
}

// ------------------------------------------------------------------ neighbour structure
template <int C> static void launch_cluster_walk(MisSim* s, int fill, cudaStream_t st) {
    const int nc = (s->n + C - 1) / C;
    const int3 cdim = make_int3(s->cell_dim[0], s->cell_dim[1], s->cell_dim[2]);
    if (s->merge_lists)
        k_cluster_merge<C><<<nblk((long long)nc * 32, 128), 128, 0, st>>>(s->x0m, s->nbr_start, s->nbr, s->n, s->d2_limit, fill, s->cl_start, s->cl, s->cl_count);
    else
        k_cluster_walk<C><<<nblk(nc, 128), 128, 0, st>>>(s->x0m, s->cell_lin_sorted, s->cell_start, s->cell_end, cdim, s->n, s->d2_limit, fill,
                                                         s->cl_start, s->cl, s->cl_count);
}
static void cluster_walk(MisSim* s, int fill, cudaStream_t st) {
    if (s->C == 1) launch_cluster_walk<1>(s, fill, st); else if (s->C == 2) launch_cluster_walk<2>(s, fill, st); else launch_cluster_walk<4>(s, fill, st);
    s->launches++;
}
// union neighbour list per cluster of C consecutive slots: count, scan, fill
static int build_cluster_lists(MisSim* s, cudaStream_t st) {
    const int nc = (s->n + s->C - 1) / s->C;
    cluster_walk(s, 0, st);
    CK_LAUNCH();
    s->launches += exclusive_scan<unsigned long long>(s->cl_count, s->cl_start, nc, s->scan_tmp, st);
    CK_LAUNCH();
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, s->cl_start + nc, sizeof total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    s->cl_total = (long long)total;
    if (s->cl_total > s->cl_cap || !s->cl) {
        if (s->cl) cudaFree(s->cl);
        s->cl = nullptr;
        CK(dalloc(&s->cl, (size_t)s->cl_total + LIST_PAD));
        CK(cudaMemsetAsync(s->cl, 0, ((size_t)s->cl_total + LIST_PAD) * sizeof(uint32_t), st));
        s->cl_cap = s->cl_total;
    }
    cluster_walk(s, 1, st);
    CK_LAUNCH();
    return MIS_OK;
}

// ---- shared-memory cell tiles: active-cell table, uint16 tile-offset lists (mis_tile.cuh).  Needs the exact lists.
static const int TILE_CAPS_D[] = {2048, 2560, 3072, 4096};
static const int TILE_CAPS_F[] = {2048, 2560, 2816};
template <int CAP, int MINB> static cudaError_t tile_attr_d() {
    return cudaFuncSetAttribute(k_deform_t<CAP, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_smem_deform<CAP>());
}
template <int CAP> static cudaError_t tile_attr_f() {
    return cudaFuncSetAttribute(k_force_t<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_smem_force<CAP>());
}
// active-cell table of the current binning: flag / scan / one row per non-empty cell; largest tile and cell
static int build_tile_tab(MisSim* s, cudaStream_t st) {
    MisSim::Tile& t = s->tile;
    const int n = s->n;
    t.ok = false; t.tab_ok = false;
    if (!t.flag) {
        CK(dalloc(&t.flag, (size_t)n)); CK(dalloc(&t.pos, (size_t)n + 1)); CK(dalloc(&t.nblocks, (size_t)n));
        CK(dalloc(&t.blk_start, (size_t)n + 1)); CK(dalloc(&t.t_off, (size_t)n)); CK(dalloc(&t.AB, 5 * (size_t)n)); CK(dalloc(&t.max_dev, 4));
        CK(dalloc(&t.wkey, (size_t)n)); CK(dalloc(&t.order, (size_t)n));
    }
    const int3 cdim = make_int3(s->cell_dim[0], s->cell_dim[1], s->cell_dim[2]);
    k_tile_flag<<<nblk(n, 256), 256, 0, st>>>(s->cell_lin_sorted, s->cell_start, n, t.flag);
    CK_LAUNCH(); s->launches++;
    s->launches += exclusive_scan<unsigned long long>(t.flag, t.pos, n, s->scan_tmp, st);
    unsigned long long na = 0;
    CK(cudaMemcpyAsync(&na, t.pos + n, sizeof na, cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(t.max_dev, 0, 2 * sizeof(int), st));
    CK(cudaStreamSynchronize(st));
    t.n_active = (int)na;
    if ((long long)t.n_active > t.tab_cap) {
        if (t.tab) cudaFree(t.tab);
        t.tab = nullptr;
        CK(dalloc(&t.tab, (size_t)t.n_active * TT_STRIDE));
        t.tab_cap = t.n_active;
    }
    const int by_cell = s->ncells <= n ? 1 : 0;
    const int items = by_cell ? s->ncells : n;
    k_tile_tab<<<nblk((long long)items * 32, 256), 256, 0, st>>>(t.flag, t.pos, s->cell_lin_sorted, s->cell_start, s->cell_end, cdim, items, by_cell, t.tab, t.max_dev);
    CK_LAUNCH(); s->launches++;
    int mx[2] = {0, 0};
    CK(cudaMemcpyAsync(mx, t.max_dev, sizeof mx, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    t.max_tile = mx[0]; t.max_own = mx[1];
    t.cap_d = t.cap_f = 0;
    for (int c : TILE_CAPS_D) if (!t.cap_d && t.max_tile <= c) t.cap_d = c;
    for (int c : TILE_CAPS_F) if (!t.cap_f && t.max_tile <= c) t.cap_f = c;
    t.W = (t.max_tile + 31) / 32;
    t.tab_ok = true;
    return MIS_OK;
}
// list blocks per particle -> offsets; allocates the uint16 lists.  Needs nbr_count.  One host round trip.
static int alloc_tile_lists(MisSim* s, cudaStream_t st) {
    MisSim::Tile& t = s->tile;
    const int n = s->n;
    k_tile_count<<<nblk(n, 256), 256, 0, st>>>(s->nbr_count, n, t.nblocks);
    CK_LAUNCH(); s->launches++;
    s->launches += exclusive_scan<unsigned long long>(t.nblocks, t.blk_start, n, s->scan_tmp, st);
    unsigned long long tb = 0;
    CK(cudaMemcpyAsync(&tb, t.blk_start + n, sizeof tb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    t.total_blocks = (long long)tb;
    if (t.total_blocks > t.lists_cap || !t.lists) {
        if (t.lists) cudaFree(t.lists);
        t.lists = nullptr;
        CK(dalloc(&t.lists, (size_t)(t.total_blocks + 1) * TILE_BLOCK));
        t.lists_cap = t.total_blocks;
    }
    return MIS_OK;
}
// launch order (longest cell first) and the per-device function attributes of the instantiations this scene uses
static int finish_tiles(MisSim* s, cudaStream_t st) {
    MisSim::Tile& t = s->tile;
    k_tile_work<<<nblk(t.n_active, 256), 256, 0, st>>>(t.tab, t.n_active, s->nbr_start, t.wkey, t.order);
    CK_LAUNCH(); s->launches++;
    s->launches += radix_sort_pairs(t.wkey, t.order, t.n_active, 16, s->rs, st);
    CK_LAUNCH();
    cudaError_t e = cudaSuccess;
    switch (t.cap_d) {
        case 2048: e = tile_attr_d<2048, 3>(); break; case 2560: e = tile_attr_d<2560, 2>(); break;
        case 3072: e = tile_attr_d<3072, 2>(); break; default: e = tile_attr_d<4096, 1>(); break;
    }
    CK(e);
    switch (t.cap_f) { case 2048: e = tile_attr_f<2048>(); break; case 2560: e = tile_attr_f<2560>(); break; default: e = tile_attr_f<2816>(); break; }
    CK(e);
    t.ok = true;
    return MIS_OK;
}
// uint16 tile lists from exact lists that the per-thread walk built (scenes whose tiles exceed the bitmask build)
static int build_tiles(MisSim* s, cudaStream_t st) {
    MisSim::Tile& t = s->tile;
    if (t.ok) return MIS_OK;
    if (!t.tab_ok) { int rc = build_tile_tab(s, st); if (rc) return rc; }
    if (!t.cap_d || !t.cap_f) return MIS_OK;          // a neighbourhood too dense for one tile: the cluster kernels run instead
    int rc = alloc_tile_lists(s, st);
    if (rc) return rc;
    const int3 cdim = make_int3(s->cell_dim[0], s->cell_dim[1], s->cell_dim[2]);
    k_tile_lists<<<nblk((long long)s->n * 32, 256), 256, 0, st>>>(s->nbr_start, s->nbr, s->nbr_count, s->cell_lin_sorted, s->cell_start, t.pos, t.tab, cdim, s->n,
                                                                 t.blk_start, t.t_off, t.lists);
    CK_LAUNCH(); s->launches++;
    return finish_tiles(s, st);
}
// Exact lists, uint16 tile lists and (clusters of 2) union lists from ONE pass of distance tests over shared-memory tiles.
static int build_lists_tiled(MisSim* s, cudaStream_t st) {
    MisSim::Tile& t = s->tile;
    const int n = s->n, W = t.W;
    const int3 cdim = make_int3(s->cell_dim[0], s->cell_dim[1], s->cell_dim[2]);
    if ((long long)n * W > t.bits_cap) {
        if (t.bits) cudaFree(t.bits);
        t.bits = nullptr;
        CK(dalloc(&t.bits, (size_t)n * W));
        t.bits_cap = (long long)n * W;
    }
    const bool pairs = s->C == 2 && !s->merge_lists;
    int* straddle = reinterpret_cast<int*>(t.wkey);                      // n ints, free until finish_tiles: the clusters k_cluster_merge works on
    int n_straddle = 0;
    if (pairs) {
        CK(cudaMemsetAsync(t.max_dev + 2, 0, sizeof(int), st));
        k_straddle_list<<<nblk((n + 1) / 2, 256), 256, 0, st>>>(s->cell_lin_sorted, n, straddle, t.max_dev + 2);
        CK_LAUNCH(); s->launches++;
        CK(cudaMemcpyAsync(&n_straddle, t.max_dev + 2, sizeof(int), cudaMemcpyDeviceToHost, st));     // on the host after alloc_tile_lists' sync
    }
    const int smem = tb_smem_bytes(W, t.max_own);
    CK(cudaFuncSetAttribute(k_tile_walk_bits, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemsetAsync(s->max_k_dev, 0, sizeof(int), st));
    k_tile_walk_bits<<<t.n_active, TB_THREADS, smem, st>>>(nullptr, t.tab, s->x0m, n, s->d2_limit, W, t.max_own, t.bits, s->nbr_count, s->max_k_dev,
                                                          pairs ? s->cl_count : nullptr);
    CK_LAUNCH(); s->launches++;
    s->launches += exclusive_scan<unsigned long long>(s->nbr_count, s->nbr_start, n, s->scan_tmp, st);
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, s->nbr_start + n, sizeof total, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&s->max_k, s->max_k_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    int rc = alloc_tile_lists(s, st);                                   // synchronises: total / max_k are on the host too
    if (rc) return rc;
    s->total_pairs = (long long)total;
    if (s->total_pairs > s->nbr_cap) {
        if (s->nbr) cudaFree(s->nbr);
        s->nbr = nullptr;
        CK(dalloc(&s->nbr, (size_t)s->total_pairs + 32));
        s->nbr_cap = s->total_pairs;
    }
    // per-cell expansion with warp staging buffers (mis_tilebuild.cuh) unless a list is longer than the staging capacity
    const bool old_expand = env_int("MIS_BUILD_EXPAND_OLD", 0) != 0;           // A/B: the one-bit-per-lane expansion (read per build)
    const bool cell_expand = s->max_k <= TX_MAX_K && !old_expand;
    const int KP = ((s->max_k > 0 ? s->max_k : 1) + TILE_BLOCK - 1) / TILE_BLOCK * TILE_BLOCK;
    if (cell_expand) {
        CK(cudaFuncSetAttribute(k_tile_expand<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tx_smem_bytes(W, KP, 0)));
        k_tile_expand<0><<<t.n_active, TX_THREADS, tx_smem_bytes(W, KP, 0), st>>>(t.bits, W, KP, t.tab, s->nbr_start, s->nbr, t.blk_start, t.t_off, t.lists);
    } else {
        k_bits_expand<0><<<nblk((long long)n * 32, 256), 256, 0, st>>>(t.bits, W, s->cell_lin_sorted, s->cell_start, t.pos, t.tab, n, s->nbr_start, s->nbr,
                                                                      s->nbr_count, t.blk_start, t.t_off, t.lists);
    }
    CK_LAUNCH(); s->launches++;
    if (pairs) {
        const int nc = (n + 1) / 2;
        // clusters that straddle a cell boundary: union from the members' exact lists (count, then fill after the scan)
        if (n_straddle > 0)
            k_cluster_merge<2><<<nblk((long long)n_straddle * 32, 128), 128, 0, st>>>(s->x0m, s->nbr_start, s->nbr, n, s->d2_limit, 0, s->cl_start, s->cl, s->cl_count,
                                                                                      s->cell_lin_sorted, straddle, n_straddle);
        CK_LAUNCH(); s->launches++;
        s->launches += exclusive_scan<unsigned long long>(s->cl_count, s->cl_start, nc, s->scan_tmp, st);
        unsigned long long ct = 0;
        CK(cudaMemcpyAsync(&ct, s->cl_start + nc, sizeof ct, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        s->cl_total = (long long)ct;
        if (s->cl_total > s->cl_cap || !s->cl) {
            if (s->cl) cudaFree(s->cl);
            s->cl = nullptr;
            CK(dalloc(&s->cl, (size_t)s->cl_total + LIST_PAD));
            s->cl_cap = s->cl_total;
        }
        CK(cudaMemsetAsync(s->cl + s->cl_total, 0, (size_t)LIST_PAD * sizeof(uint32_t), st));      // the zero entries the force kernel may read past the last list
        if (cell_expand) {
            CK(cudaFuncSetAttribute(k_tile_expand<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tx_smem_bytes(W, KP, 1)));
            k_tile_expand<1><<<t.n_active, TX_THREADS, tx_smem_bytes(W, KP, 1), st>>>(t.bits, W, KP, t.tab, s->cl_start, s->cl, nullptr, nullptr, nullptr);
        } else {
            k_bits_expand<1><<<nblk((long long)nc * 32, 256), 256, 0, st>>>(t.bits, W, s->cell_lin_sorted, s->cell_start, t.pos, t.tab, n, s->cl_start, s->cl,
                                                                           nullptr, nullptr, nullptr, nullptr);
        }
        if (n_straddle > 0)
            k_cluster_merge<2><<<nblk((long long)n_straddle * 32, 128), 128, 0, st>>>(s->x0m, s->nbr_start, s->nbr, n, s->d2_limit, 1, s->cl_start, s->cl, s->cl_count,
                                                                                      s->cell_lin_sorted, straddle, n_straddle);
        CK_LAUNCH(); s->launches += 2;
    } else {
        rc = build_cluster_lists(s, st);
        if (rc) return rc;
    }
    if (t.cap_d && t.cap_f) return finish_tiles(s, st);
    return MIS_OK;
}
static TileView make_tile_view(MisSim* s) {
    TileView t;
    t.order = s->tile.order; t.tab = s->tile.tab; t.n_active = s->tile.n_active; t.lists = (const uint4*)s->tile.lists;
    t.t_off = s->tile.t_off; t.t_cnt = s->nbr_count; t.AB = s->tile.AB;
    return t;
}
static bool use_tiles_d(const MisSim* s) { return s->tile.want_d && s->tile.ok && !s->p.two_pass_deform; }
static bool use_tiles_f(const MisSim* s) { return s->tile.want_f && s->tile.ok && !s->p.symmetric_pair; }

extern "C" int mis_build_neighbors(MisSim* s, void* stream) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    if (s->ref64) return fail(MIS_E_UNSUPPORTED, "fp64 scenes keep the neighbour structure of mis_create (a rebuild is idempotent anyway: queries are on x0)");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = s->n;
    const float cw = 2.f * s->p.h;            // real(2.) * h, sim.py:127
    const float inv_cw = 1.f / cw;
    int hb[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    CK(cudaMemcpyAsync(s->bounds_dev, hb, sizeof hb, cudaMemcpyHostToDevice, st));
    k_cell_coords<<<nblk(n, 256), 256, 0, st>>>(s->x0_orig, n, inv_cw, s->p.grid_x, s->p.grid_y, s->p.grid_z,
                                                s->coords, s->bcoords, s->cell_index, s->subkey, s->bounds_dev);
    CK_LAUNCH(); s->launches++;
    CK(cudaMemcpyAsync(hb, s->bounds_dev, sizeof hb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int maxdim = 0;
    for (int a = 0; a < 3; a++) {
        s->cell_min[a] = hb[a];
        s->cell_dim[a] = hb[3 + a] - hb[a] + 1;
        if (s->cell_dim[a] > maxdim) maxdim = s->cell_dim[a];
    }
    long long nc = (long long)s->cell_dim[0] * s->cell_dim[1] * s->cell_dim[2];
    if (nc > (1ll << 28)) return fail(MIS_E_UNSUPPORTED, "dense cell table larger than 2^28 cells (the bounding box of x0 in cells of width 2h)");
    s->ncells = (int)nc;
    if (s->ncells > s->ncells_cap) {
        if (s->cell_start) cudaFree(s->cell_start);
        if (s->cell_end) cudaFree(s->cell_end);
        s->cell_start = s->cell_end = nullptr;
        CK(dalloc(&s->cell_start, (size_t)s->ncells));
        CK(dalloc(&s->cell_end, (size_t)s->ncells));
        s->ncells_cap = s->ncells;
    }
    CK(cudaMemsetAsync(s->cell_start, 0, (size_t)s->ncells * sizeof(int), st));
    CK(cudaMemsetAsync(s->cell_end, 0, (size_t)s->ncells * sizeof(int), st));
    const int3 cmin = make_int3(s->cell_min[0], s->cell_min[1], s->cell_min[2]);
    const int3 cdim = make_int3(s->cell_dim[0], s->cell_dim[1], s->cell_dim[2]);
    int bits = 1;
    while ((1 << bits) < maxdim) bits++;
    // Up to 1024 cells per axis: the plain 3 x bits Morton code.  Wider scenes: a bit count per axis (morton3_axes); their sum is
    // at most 31 because the dense table holds at most 2^28 cells.
    int3 nbits = make_int3(bits, bits, bits);
    int cell_bits = 3 * bits;
    if (maxdim > 1024) {
        int ab[3];
        for (int a = 0; a < 3; a++) { ab[a] = 0; while ((1 << ab[a]) < s->cell_dim[a]) ab[a]++; }
        nbits = make_int3(ab[0], ab[1], ab[2]);
        cell_bits = ab[0] + ab[1] + ab[2];
        if (cell_bits > 32) return fail(MIS_E_UNSUPPORTED, "cell grid needs more than 32 Morton key bits");
    }
    // in-cell Morton refinement: as many of its 9 bits as fit a 32-bit key (multiples of 3)
    s->sub_bits = cell_bits + 9 <= 32 ? 9 : (cell_bits + 6 <= 32 ? 6 : (cell_bits + 3 <= 32 ? 3 : 0));
    k_cell_keys<<<nblk(n, 256), 256, 0, st>>>(s->bcoords, s->subkey, n, cmin, nbits, s->sub_bits, s->keys);
    CK_LAUNCH(); s->launches++;
    s->launches += radix_sort_pairs(s->keys, s->perm, n, cell_bits + s->sub_bits, s->rs, st);
    CK_LAUNCH();
    k_cell_table<<<nblk(n, 256), 256, 0, st>>>(s->keys, s->perm, s->bcoords, s->x0_orig, n, cmin, cdim, s->sub_bits,
                                               s->cell_start, s->cell_end, s->cell_lin_sorted, s->inv_perm, s->x0m);
    CK_LAUNCH(); s->launches++;
    if (s->C > 1 && s->pair_cells) {
        // close particles into the same cluster: greedy nearest-neighbour matching inside every cell (mis_neighbors.cuh)
        k_pair_cells<<<nblk(s->ncells, 4), 128, 0, st>>>(s->x0m, s->cell_start, s->cell_end, s->ncells, s->perm, s->rs.vals_alt);
        CK(cudaMemcpyAsync(s->perm, s->rs.vals_alt, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        k_apply_order<<<nblk(n, 256), 256, 0, st>>>(s->perm, s->x0_orig, n, s->inv_perm, s->x0m);
        CK_LAUNCH(); s->launches += 2;
    }
    s->tile.ok = false; s->tile.tab_ok = false;
    int rc = build_tile_tab(s, st);
    if (rc) return rc;
    static const bool force_walk = env_int("MIS_BUILD_WALK", 0) != 0;          // A/B: the per-thread 27-cell walks of round 1
    if (s->tile.max_tile <= 32 * TB_MAX_WORDS && tb_smem_bytes(s->tile.W, s->tile.max_own) <= 227 * 1024 && !force_walk) {
        rc = build_lists_tiled(s, st);
        if (rc) return rc;
    } else {
        CK(cudaMemsetAsync(s->max_k_dev, 0, sizeof(int), st));
        k_nbr_walk<<<nblk(n, 128), 128, 0, st>>>(s->x0m, s->cell_lin_sorted, s->cell_start, s->cell_end, cdim, n, s->d2_limit, 0,
                                                 nullptr, nullptr, s->nbr_count, s->max_k_dev);
        CK_LAUNCH(); s->launches++;
        s->launches += exclusive_scan<unsigned long long>(s->nbr_count, s->nbr_start, n, s->scan_tmp, st);
        CK_LAUNCH();
        unsigned long long total = 0;
        CK(cudaMemcpyAsync(&total, s->nbr_start + n, sizeof total, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&s->max_k, s->max_k_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        s->total_pairs = (long long)total;
        if (s->total_pairs > s->nbr_cap) {
            if (s->nbr) cudaFree(s->nbr);
            s->nbr = nullptr;
            CK(dalloc(&s->nbr, (size_t)s->total_pairs + 32));
            s->nbr_cap = s->total_pairs;
        }
        k_nbr_walk<<<nblk(n, 128), 128, 0, st>>>(s->x0m, s->cell_lin_sorted, s->cell_start, s->cell_end, cdim, n, s->d2_limit, 1,
                                                 s->nbr_start, s->nbr, s->nbr_count, s->max_k_dev);
        CK_LAUNCH(); s->launches++;
        rc = build_cluster_lists(s, st);
        if (rc) return rc;
        if (s->tile.want_d || s->tile.want_f) { rc = build_tiles(s, st); if (rc) return rc; }
    }
    {   // equal-work cluster ranges per SM for the persistent force kernel
        if (!s->nsm) {
            int dev = 0;
            CK(cudaGetDevice(&dev));
            CK(cudaDeviceGetAttribute(&s->nsm, cudaDevAttrMultiProcessorCount, dev));
            CK(dalloc(&s->sm_first, (size_t)s->nsm + 1));
            CK(dalloc(&s->sm_ctr, (size_t)s->nsm));
        }
        const int ncl = (n + s->C - 1) / s->C;
        k_sm_ranges<<<nblk(s->nsm + 1, 256), 256, 0, st>>>(s->cl_start, ncl, s->nsm, s->sm_first);
        CK_LAUNCH(); s->launches++;
    }
    drop_graph(s);                            // a captured chunk holds the old list pointers
    s->built = true;
    return MIS_OK;
}


// ------------------------------------------------------------------ precision-templated engine (mis_ref.cuh)
// statics of an engine from the fp32 scene state: x0, m, fext, free_points from the float4 arrays, E / nu / design from raw_in
template <typename T> static void ref_pull(MisSim* s, RefEngine<T>& g, unsigned what, cudaStream_t st) {
    const int n = s->n, b = nblk(n, 256);
    g.s.nbr_start = s->nbr_start; g.s.nbr = s->nbr;
    if (what & 1u) ref::kr_pull<T><<<b, 256, 0, st>>>(s->x0m, n, 0, 3, g.s.x0);
    if (what & 2u) ref::kr_pull<T><<<b, 256, 0, st>>>(s->x0m, n, 3, 1, g.s.m);
    if (what & 4u) {
        ref::kr_cast<T, float><<<b, 256, 0, st>>>(s->raw_in, n, g.E);
        ref::kr_cast<T, float><<<b, 256, 0, st>>>(s->raw_in + n, n, g.nu);
    }
    if (what & 8u) ref::kr_cast<T, float><<<b, 256, 0, st>>>(s->raw_in + 2 * (size_t)n, n, g.design);
    if (what & 16u) ref::kr_pull<T><<<b, 256, 0, st>>>(s->fext, n, 0, 3, g.s.fext);
    if (what & 32u) ref::kr_pull<T><<<b, 256, 0, st>>>(s->freem, n, 0, 3, g.s.freem);
    if (what & (2u | 4u | 8u)) g.statics_dirty = true;
    g.primed = false;
}
static int ref64_attach(MisSim* s, cudaStream_t st) {
    if (s->p.two_pass_deform) return fail(MIS_E_UNSUPPORTED, "fp64: two_pass_deform is an fp32 tuning flag");
    RefEngine<double>* g = new RefEngine<double>();
    s->ref64 = g;
    CK(ref_create(*g, s->n, s->nbr_start, s->nbr, s->p));
    ref_pull(s, *g, 1u | 32u, st);
    CK_LAUNCH();
    return MIS_OK;
}
// the float setters of an fp64 scene forward their (exactly converted) values to the engine
static void ref64_after_set(MisSim* s, unsigned what, cudaStream_t st) { if (s->ref64) ref_pull(s, *R64(s), what, st); }

template <typename T> static T* f64_target(RefEngine<T>& g, int what, int* dim, bool* is_static) {
    *is_static = true;
    switch (what) {
        case F64_X0: *dim = 3; return g.s.x0;          case F64_MASS: *dim = 1; return g.s.m;
        case F64_YOUNGS: *dim = 1; return g.E;         case F64_POISSON: *dim = 1; return g.nu;
        case F64_DESIGN: *dim = 1; return g.design;    case F64_EXT_FORCE: *dim = 3; return g.s.fext;
        case F64_DIRICHLET: *dim = 3; return g.s.freem;
        case F64_POSITION: *dim = 3; *is_static = false; return g.s.x;
        case F64_VELOCITY: *dim = 3; *is_static = false; return g.s.v;
        case F64_ELASTIC_FORCE: *dim = 3; *is_static = false; return g.s.fel;
        case F64_VOLUME: *dim = 1; return g.s.vol;     case F64_RHO: *dim = 1; return g.s.rho;
        case F64_DEF_GRAD: *dim = 9; *is_static = false; return g.s.F;   case F64_STRESS: *dim = 9; *is_static = false; return g.s.S;
        case F64_ROTATION: *dim = 9; *is_static = false; return g.s.R;   case F64_A_PQ: *dim = 9; *is_static = false; return g.s.A;
    }
    return nullptr;
}

extern "C" int mis_set_constants_f64(MisSim* s, const double c[8]) {
    if (!s || !c) return fail(MIS_E_INVALID, "null argument");
    if (!s->ref64) return fail(MIS_E_STATE, "mis_set_constants_f64 needs a scene created with MisParams.fp64 = 1");
    if (!(c[0] > 0.0) || !(c[2] > 0.0)) return fail(MIS_E_INVALID, "h and dt must be positive");
    ref::RP<double>& r = R64(s)->c;
    r.h = c[0]; r.damping = c[1]; r.dt = c[2]; r.k_col = c[3]; r.col_range = c[4]; r.stiff_a = c[5]; r.stiff_b = c[6]; r.tanh_k = c[7];
    R64(s)->statics_dirty = true; R64(s)->primed = false;
    return MIS_OK;
}

extern "C" int mis_set_f64(MisSim* s, int what, const double* src_dev, void* stream) {
    if (!s || !src_dev) return fail(MIS_E_INVALID, "null argument");
    if (!s->ref64) return fail(MIS_E_STATE, "mis_set_f64 needs a scene created with MisParams.fp64 = 1");
    if (what < F64_X0 || what > F64_VELOCITY) return fail(MIS_E_INVALID, "mis_set_f64: `what` must be one of x0, mass, youngs, poisson, design, ext_force, dirichlet, position, velocity");
    cudaStream_t st = (cudaStream_t)stream;
    RefEngine<double>& g = *R64(s);
    int dim = 0; bool stat = false;
    double* dst = f64_target(g, what, &dim, &stat);
    ref::kr_gather<double, double><<<nblk(s->n, 256), 256, 0, st>>>(src_dev, s->perm, s->n, dim, dst);
    CK_LAUNCH(); s->launches++;
    if (what == F64_X0 || what == F64_MASS || what == F64_YOUNGS || what == F64_POISSON || what == F64_DESIGN) g.statics_dirty = true;
    if (what == F64_MASS) s->mass_set = true;
    if (what == F64_YOUNGS || what == F64_POISSON) s->material_set = true;
    if (what == F64_POSITION || what == F64_VELOCITY) s->started = true;
    g.primed = false;
    return MIS_OK;
}

// engine state the getters read must describe the current frame: statics computed, fields evaluated at x
static void ref64_refresh(MisSim* s, cudaStream_t st) {
    RefEngine<double>& g = *R64(s);
    ref_statics(g, st);
    if (s->started && !(g.primed && !g.c.euler)) { ref_eval(g, g.s.x, g.s.fel, st); g.primed = !g.c.euler; }
}

extern "C" int mis_get_f64(MisSim* s, int what, double* dst_dev, void* stream) {
    if (!s || !dst_dev) return fail(MIS_E_INVALID, "null argument");
    if (!s->ref64) return fail(MIS_E_STATE, "mis_get_f64 needs a scene created with MisParams.fp64 = 1");
    cudaStream_t st = (cudaStream_t)stream;
    RefEngine<double>& g = *R64(s);
    int dim = 0; bool stat = false;
    double* src = f64_target(g, what, &dim, &stat);
    if (!src) return fail(MIS_E_INVALID, "mis_get_f64: unknown `what`");
    if (what >= F64_ELASTIC_FORCE) ref64_refresh(s, st);
    ref::kr_scatter<double, double><<<nblk(s->n, 256), 256, 0, st>>>(src, s->perm, s->n, dim, dst_dev);
    CK_LAUNCH(); s->launches++;
    return MIS_OK;
}

// Reverse pass of the rollout (sim.py:341-372, compute_grad=True): loss value and d loss / d design x.
extern "C" int mis_rollout_grad(MisSim* s, int frames, int n_targets, const float* target_x_dev, const float* target_v_dev,
                                int checkpoint_every, double* loss_host, float* grad_dev, double* grad64_dev, void* stream) {
    if (!s || frames <= 0 || n_targets < 0 || !loss_host || (n_targets > 0 && (!target_x_dev || !target_v_dev)))
        return fail(MIS_E_INVALID, "mis_rollout_grad: bad argument");
    if (!s->mass_set || !s->material_set) return fail(MIS_E_STATE, "set_mass and set_material must precede mis_rollout_grad");
    if (s->p.euler) return fail(MIS_E_UNSUPPORTED, "mis_rollout_grad differentiates the velocity-Verlet loop of sim.py (not the Euler variant)");
    if (s->sdf || s->halo_on) return fail(MIS_E_UNSUPPORTED, "mis_rollout_grad: obstacle contact and slab partitions are outside the differentiated path");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = s->n;
    if (checkpoint_every <= 0) { checkpoint_every = 1; while (checkpoint_every * checkpoint_every < frames) checkpoint_every++; }
    cudaError_t e;
    if (s->ref64) {
        RefEngine<double>& g = *R64(s);
        e = ref_rollout_grad(g, s->v0, frames, n_targets, target_x_dev, target_v_dev, s->perm, checkpoint_every, loss_host, st);
        CK(e);
        ref::kr_design_grad<double><<<nblk(n, 256), 256, 0, st>>>(n, g.ratio_b, g.design, g.c.tanh_k, s->perm, grad_dev, grad64_dev);
    } else {
        if (!s->ref32) {
            RefEngine<float>* g = new RefEngine<float>();
            s->ref32 = g;
            CK(ref_create(*g, n, s->nbr_start, s->nbr, s->p));
        }
        RefEngine<float>& g = *R32(s);
        ref_pull(s, g, 63u, st);
        e = ref_rollout_grad(g, s->v0, frames, n_targets, target_x_dev, target_v_dev, s->perm, checkpoint_every, loss_host, st);
        CK(e);
        ref::kr_design_grad<float><<<nblk(n, 256), 256, 0, st>>>(n, g.ratio_b, g.design, g.c.tanh_k, s->perm, grad_dev, grad64_dev);
    }
    CK_LAUNCH();
    CK(cudaStreamSynchronize(st));
    s->launches += 8LL * frames;
    s->dirty = true; s->forces_only = false;
    return MIS_OK;
}

extern "C" int mis_get_neighbor_info(MisSim* s, MisNeighborInfo* out) {
    if (!s || !out) return fail(MIS_E_INVALID, "null argument");
    out->total_pairs = s->total_pairs; out->max_neighbors = s->max_k; out->n = s->n;
    for (int a = 0; a < 3; a++) { out->cell_min[a] = s->cell_min[a]; out->cell_dim[a] = s->cell_dim[a]; }
    out->cell_width = 2.f * s->p.h;
    out->cluster_size = s->C; out->union_entries = s->cl_total;
    return MIS_OK;
}

extern "C" int mis_export_cells(MisSim* s, int* cell_index_dev, int* cell_coords_dev, int* perm_dev, void* stream) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)s->n;
    if (cell_index_dev) CK(cudaMemcpyAsync(cell_index_dev, s->cell_index, N * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (cell_coords_dev) CK(cudaMemcpyAsync(cell_coords_dev, s->coords, 3 * N * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (perm_dev) CK(cudaMemcpyAsync(perm_dev, s->perm, N * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return MIS_OK;
}

extern "C" int mis_export_cell_ranges(MisSim* s, int* start_dev, int* end_dev, void* stream) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    cudaStream_t st = (cudaStream_t)stream;
    if (start_dev) CK(cudaMemcpyAsync(start_dev, s->cell_start, (size_t)s->ncells * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (end_dev) CK(cudaMemcpyAsync(end_dev, s->cell_end, (size_t)s->ncells * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return MIS_OK;
}

extern "C" int mis_export_neighbors(MisSim* s, long long* offsets_dev, int* nbr_dev, void* stream) {
    if (!s || !offsets_dev) return fail(MIS_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = s->n;
    uint32_t* counts = nullptr;
    unsigned long long* tmp = nullptr;
    CK(dalloc(&counts, (size_t)n));
    CK(dalloc(&tmp, (size_t)nblk(n, SCAN_TILE) + 1));
    k_export_counts<<<nblk(n, 256), 256, 0, st>>>(s->nbr_count, s->inv_perm, n, counts);
    s->launches += 1 + exclusive_scan<unsigned long long>(counts, (unsigned long long*)offsets_dev, n, tmp, st);
    if (nbr_dev) {
        k_export_lists<<<nblk((long long)n * 32, 256), 256, 0, st>>>(s->nbr_start, s->nbr, s->inv_perm, s->perm, n, offsets_dev, nbr_dev);
        s->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(counts); cudaFree(tmp);
    CK(e);
    CK_LAUNCH();
    return MIS_OK;
}

// ------------------------------------------------------------------ control functions
template <int G> static void launch_volume(MisSim* s, cudaStream_t st) {
    k_volume<G><<<nblk((long long)s->n * G, STEP_THREADS), STEP_THREADS, 0, st>>>(s->x0m, s->nbr_start, s->nbr, s->n, s->c,
                                                                                  s->p.self_density, s->xv[0], s->xv[1], s->matl);
    k_static_K<G><<<nblk((long long)s->n * G, STEP_THREADS), STEP_THREADS, 0, st>>>(s->x0m, s->xv[0], s->nbr_start, s->nbr, s->n, s->c, s->Ks);
}

extern "C" int mis_set_mass(MisSim* s, const float* mass_dev, void* stream) {
    if (!s || !mass_dev) return fail(MIS_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    k_gather_w<<<nblk(s->n, 256), 256, 0, st>>>(mass_dev, s->perm, s->n, s->x0m);
    CK_LAUNCH();
    if (s->G == 8) launch_volume<8>(s, st); else if (s->G == 16) launch_volume<16>(s, st); else launch_volume<32>(s, st);
    CK_LAUNCH();
    s->launches += 3;
    s->mass_set = true; s->dirty = true; s->forces_only = false;
    ref64_after_set(s, 2u, st);
    return MIS_OK;
}

extern "C" int mis_set_material(MisSim* s, const float* youngs_dev, const float* poisson_dev, void* stream) {
    if (!s || !youngs_dev || !poisson_dev) return fail(MIS_E_INVALID, "null argument");
    k_material<<<nblk(s->n, 256), 256, 0, (cudaStream_t)stream>>>(youngs_dev, poisson_dev, s->perm, s->n, s->matl);
    ref::kr_gather<float, float><<<nblk(s->n, 256), 256, 0, (cudaStream_t)stream>>>(youngs_dev, s->perm, s->n, 1, s->raw_in);
    ref::kr_gather<float, float><<<nblk(s->n, 256), 256, 0, (cudaStream_t)stream>>>(poisson_dev, s->perm, s->n, 1, s->raw_in + s->n);
    CK_LAUNCH(); s->launches += 3;
    s->material_set = true; s->dirty = true; s->forces_only = false;
    ref64_after_set(s, 4u, (cudaStream_t)stream);
    return MIS_OK;
}

extern "C" int mis_set_design(MisSim* s, const float* x_dev, void* stream) {
    if (!s || !x_dev) return fail(MIS_E_INVALID, "null argument");
    k_design<<<nblk(s->n, 256), 256, 0, (cudaStream_t)stream>>>(x_dev, s->perm, s->n, s->p.tanh_k, s->matl);
    ref::kr_gather<float, float><<<nblk(s->n, 256), 256, 0, (cudaStream_t)stream>>>(x_dev, s->perm, s->n, 1, s->raw_in + 2 * (size_t)s->n);
    CK_LAUNCH(); s->launches += 2;
    s->dirty = true; s->forces_only = false;
    ref64_after_set(s, 8u, (cudaStream_t)stream);
    return MIS_OK;
}

extern "C" int mis_set_ext_force(MisSim* s, const float* f_dev, void* stream) {
    if (!s || !f_dev) return fail(MIS_E_INVALID, "null argument");
    k_gather_vec3<<<nblk(s->n, 256), 256, 0, (cudaStream_t)stream>>>(f_dev, s->perm, s->n, s->fext, 0);
    CK_LAUNCH(); s->launches++;
    if (!s->dirty) { s->dirty = true; s->forces_only = true; }
    ref64_after_set(s, 16u, (cudaStream_t)stream);
    return MIS_OK;
}

extern "C" int mis_set_ext_force_host(MisSim* s, const float* f_host, void* stream) {
    NO_FP64("host streaming");
    if (!s || !f_host) return fail(MIS_E_INVALID, "null argument");
    // upload on the copy stream (overlaps the kernels still running on `stream`), then gather on `stream`
    int rc = ensure_copy_stream(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int k = (int)(s->up_i++ & 1);
    CK(cudaStreamWaitEvent(s->up_stream, s->ev_up_used[k], 0));         // the gather that last read this staging slot has run
    CK(cudaMemcpyAsync(s->fstage[k], f_host, 3 * (size_t)s->n * sizeof(float), cudaMemcpyHostToDevice, s->up_stream));
    CK(cudaEventRecord(s->ev_up[k], s->up_stream));
    CK(cudaStreamWaitEvent(st, s->ev_up[k], 0));
    rc = mis_set_ext_force(s, s->fstage[k], stream);
    if (rc) return rc;
    CK(cudaEventRecord(s->ev_up_used[k], st));
    return MIS_OK;
}

extern "C" int mis_set_dirichlet(MisSim* s, const float* free_dev, void* stream) {
    if (!s || !free_dev) return fail(MIS_E_INVALID, "null argument");
    k_gather_vec3<<<nblk(s->n, 256), 256, 0, (cudaStream_t)stream>>>(free_dev, s->perm, s->n, s->freem, 0);
    CK_LAUNCH(); s->launches++;
    if (!s->dirty) { s->dirty = true; s->forces_only = true; }
    ref64_after_set(s, 32u, (cudaStream_t)stream);
    return MIS_OK;
}

// ------------------------------------------------------------------ step machinery
template <int C, int G> static void launch_deform(MisSim* s, const View& v, cudaStream_t st) {
    const int nc = (s->n + C - 1) / C;
    static int carve_dev[64];                 // tuning hook: MIS_DEFORM_CARVEOUT = preferred shared-memory carve-out in percent (a per-device attribute)
    static bool carve_init[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    dev_ &= 63;
    if (!carve_init[dev_]) {
        carve_init[dev_] = true;
        carve_dev[dev_] = env_int("MIS_DEFORM_CARVEOUT", -1);
        if (carve_dev[dev_] >= 0) cudaFuncSetAttribute(k_deform_c<C, G, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve_dev[dev_]);
    }
    k_deform_c<C, G, false><<<nblk((long long)nc * G, STEP_THREADS), STEP_THREADS, 0, st>>>(v, s->c);
}
template <int C> static void launch_deform2(MisSim* s, const View& v, cudaStream_t st) {
    const int nc = (s->n + C - 1) / C;
    k_deform_c<C, 8, true><<<nblk((long long)nc * 8, STEP_THREADS), STEP_THREADS, 0, st>>>(v, s->c);
}
template <int C, int G> static void launch_force(MisSim* s, const View& v, int mode, cudaStream_t st) {
    const int nc = (s->n + C - 1) / C;
    k_force_c<C, G, false><<<nblk((long long)nc * G, STEP_THREADS), STEP_THREADS, 0, st>>>(v, s->c, mode);
}
template <int C, int G> static void launch_force_p(MisSim* s, const View& v, int mode, cudaStream_t st) {
    if (!s->force_bps) {
        int bps = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_force_p<C, G, false>, STEP_THREADS, 0) != cudaSuccess || bps < 1) bps = 1;
        s->force_bps = bps;
    }
    SmQueue q;
    q.first = s->sm_first; q.ctr = s->sm_ctr; q.nsm = s->nsm;
    cudaMemsetAsync(s->sm_ctr, 0, (size_t)s->nsm * sizeof(int), st);
    k_force_p<C, G, false><<<s->nsm * s->force_bps, STEP_THREADS, 0, st>>>(v, s->c, mode, q);
}
template <int C> static void launch_force_sym(MisSim* s, const View& v, int mode, cudaStream_t st) {
    const int nc = (s->n + C - 1) / C;
    k_force_c<C, 8, true><<<nblk((long long)nc * 8, STEP_THREADS), STEP_THREADS, 0, st>>>(v, s->c, mode);
}
#define MIS_DISPATCH_CG(FN, ...)                                                                  \
    do {                                                                                          \
        const int key_ = s->C * 100 + s->G;                                                       \
        switch (key_) {                                                                           \
            case 108: FN<1, 8>(__VA_ARGS__); break;   case 116: FN<1, 16>(__VA_ARGS__); break;    \
            case 132: FN<1, 32>(__VA_ARGS__); break;  case 208: FN<2, 8>(__VA_ARGS__); break;     \
            case 216: FN<2, 16>(__VA_ARGS__); break;  case 232: FN<2, 32>(__VA_ARGS__); break;    \
            case 408: FN<4, 8>(__VA_ARGS__); break;   case 416: FN<4, 16>(__VA_ARGS__); break;    \
            default: FN<4, 32>(__VA_ARGS__); break;                                               \
        }                                                                                         \
    } while (0)
static void enqueue_deform_tile(MisSim* s, const View& v, cudaStream_t st) {
    const TileView t = make_tile_view(s);
    const int g = s->tile.n_active;
    switch (s->tile.cap_d) {
        case 2048: k_deform_t<2048, 3><<<g, TILE_THREADS_D, tile_smem_deform<2048>(), st>>>(v, s->c, t); break;
        case 2560: k_deform_t<2560, 2><<<g, TILE_THREADS_D, tile_smem_deform<2560>(), st>>>(v, s->c, t); break;
        case 3072: k_deform_t<3072, 2><<<g, TILE_THREADS_D, tile_smem_deform<3072>(), st>>>(v, s->c, t); break;
        default:   k_deform_t<4096, 1><<<g, TILE_THREADS_D, tile_smem_deform<4096>(), st>>>(v, s->c, t); break;
    }
    k_deform_fin<<<nblk(s->n, 128), 128, 0, st>>>(v, s->c, s->tile.AB);
    s->launches += 2;
}
static void enqueue_force_tile(MisSim* s, const View& v, int mode, cudaStream_t st) {
    const TileView t = make_tile_view(s);
    const int g = s->tile.n_active;
    switch (s->tile.cap_f) {
        case 2048: k_force_t<2048><<<g, TILE_THREADS_F, tile_smem_force<2048>(), st>>>(v, s->c, t, mode); break;
        case 2560: k_force_t<2560><<<g, TILE_THREADS_F, tile_smem_force<2560>(), st>>>(v, s->c, t, mode); break;
        default:   k_force_t<2816><<<g, TILE_THREADS_F, tile_smem_force<2816>(), st>>>(v, s->c, t, mode); break;
    }
    s->launches++;
}
static void enqueue_deform(MisSim* s, const View& v, cudaStream_t st) {
    if (use_tiles_d(s)) { enqueue_deform_tile(s, v, st); return; }
    if (s->p.two_pass_deform) {       // reference two-loop order: accuracy studies, fixed 8 lanes per cluster
        if (s->C == 1) launch_deform2<1>(s, v, st); else if (s->C == 2) launch_deform2<2>(s, v, st); else launch_deform2<4>(s, v, st);
    } else {
        MIS_DISPATCH_CG(launch_deform, s, v, st);
    }
    s->launches++;
}
static void enqueue_force(MisSim* s, const View& v, int mode, cudaStream_t st) {
    if (use_tiles_f(s)) { enqueue_force_tile(s, v, mode, st); return; }
    if (s->p.symmetric_pair) {        // sim_taichi.py pair force, fixed 8 lanes per cluster
        if (s->C == 1) launch_force_sym<1>(s, v, mode, st); else if (s->C == 2) launch_force_sym<2>(s, v, mode, st); else launch_force_sym<4>(s, v, mode, st);
    } else if (s->force_persist && s->sm_first && (long long)(s->n / s->C) >= 1500LL * s->nsm) {
        // per-SM cluster queues pay off once an SM's range is long (measured on B200: -4.5 % at n = 1e6, +10 % at n = 1e5, where a
        // range is 340 clusters and the stealing tail shows)
        MIS_DISPATCH_CG(launch_force_p, s, v, mode, st);
    } else {
        MIS_DISPATCH_CG(launch_force, s, v, mode, st);
    }
    s->launches++;
}

// after a kernel that pushed new positions to the peers: publish this rank's epoch, wait for the peers'
static void enqueue_halo_sync(MisSim* s, cudaStream_t st) {
    if (!s->halo_on) return;
    HaloSync h;
    h.n_peers = s->halo_peers;
    h.wait = s->halo_wait ? 1 : 0;
    h.epoch = s->halo_mem + MIS_MAX_PEERS;
    h.my_flags = s->halo_mem;
    for (int p = 0; p < 4; p++) h.peer_flag[p] = s->peer_flag[p];
    h.err = (int*)(s->halo_mem + MIS_MAX_PEERS + 1);
    h.timeout_ns = s->halo_timeout_ns;
    k_halo_sync<<<1, 32, 0, st>>>(h);
    s->launches++; s->exchanges++;
}

// obstacle contact at the now-current positions: broad phase (bounding box) -> MLP values -> narrow phase (contact band)
// -> three forward-difference evaluations of the particles in contact -> penalty force.  No host synchronisation:
// the live row counts stay on the device and dead row-blocks of the GEMM grid exit at once.
static cudaError_t enqueue_contact(MisSim* s, const View& v, cudaStream_t st, cudaEvent_t after_layer0 = nullptr) {
    if (!s->sdf) return cudaSuccess;
    const int n = s->n, cap = s->con_cap;
    MisSdf* net = s->sdf;
    cudaError_t e = cudaMemsetAsync(s->con_count, 0, 3 * sizeof(int), st);        // [3] (overflow) is sticky until read
    if (e != cudaSuccess) return e;
    k_contact_select<<<nblk(n, 256), 256, 0, st>>>(v.xcur, n, s->sdf_xf, s->sdf_lo, s->sdf_hi, cap, s->con_idx, s->con_count, s->con_pts, s->fcon);
    const long long l0 = net->launches;
    int fb = 0;
    const int H = net->H;
    e = sdf_forward(net, s->con_pts, nullptr, cap, s->con_count, s->sdf_xf, make_float3(0.f, 0.f, 0.f), nullptr, st, 0, &fb, false, after_layer0);
    if (e != cudaSuccess) return e;
    e = launch_k(k_contact_last_narrow, dim3(148 * 2), dim3(256), 0, st, true,
                 (const float*)net->act[fb][0], (const float*)net->act[fb][1], cap, (const int*)s->con_count, (const float*)net->wl, (const float*)net->bl, H, s->p.col_range,
                 (const int*)s->con_idx, (const float*)s->con_pts, s->con_idx2, s->con_pts2, s->con_s0, s->con_count + 1);
    if (e != cudaSuccess) return e;
    // the three forward differences of the in-contact particles as ONE pass of 3 x count rows (row 3 r + axis): at most cap / 3
    // particles in the band at once; beyond that the narrow phase raises the overflow flag
    const float ep = s->sdf_eps;
    e = sdf_forward(net, s->con_pts2, nullptr, cap, s->con_count + 2, s->sdf_xf, make_float3(ep, 0.f, 0.f), nullptr, st, 1, &fb, true);
    if (e != cudaSuccess) return e;
    e = launch_k(k_contact_last_apply, dim3(148 * 2), dim3(256), 0, st, true,
                 (const float*)net->act[fb][0], (const float*)net->act[fb][1], cap, (const int*)(s->con_count + 1), (const float*)net->wl, (const float*)net->bl, H,
                 (const float*)s->con_s0, (const int*)s->con_idx2, 1.f / ep, s->sdf_xf, s->p.col_range, s->p.k_col, s->fcon);
    if (e != cudaSuccess) return e;
    s->launches += 3 + (net->launches - l0);
    return cudaGetLastError();
}

// k_deform_c and the contact chain of one frame: both depend on the new positions only
static void enqueue_deform_contact(MisSim* s, const View& v, cudaStream_t st) {
    cudaError_t e;
    if (s->sdf && s->side_stream && !s->serial_contact) {
        cudaEventRecord(s->ev_fork, st);
        cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0);
        // The chain's CTAs need ~200 KB of shared memory each (one per SM, 64 SMs); a deformation kernel that is already
        // resident everywhere starves them until it drains.  Launched just behind the chain's first layer instead, the chain
        // (higher-priority stream) takes its 64 SMs first and the deformation CTAs fill the other 84: the two overlap.
        e = enqueue_contact(s, v, s->side_stream, s->contact_first ? s->ev_chain : nullptr);
        cudaEventRecord(s->ev_join, s->side_stream);
        if (s->contact_first) cudaStreamWaitEvent(st, s->ev_chain, 0);
        enqueue_deform(s, v, st);
        cudaStreamWaitEvent(st, s->ev_join, 0);
    } else {
        enqueue_deform(s, v, st);
        e = enqueue_contact(s, v, st);
    }
    if (e != cudaSuccess && s->enqueue_err == cudaSuccess) s->enqueue_err = e;      // surfaced by mis_step / prime
}

// deformation + contact + force/integrate of one frame at the current positions
static void enqueue_frame(MisSim* s, const View& v, int mode, cudaStream_t st) {
    if (s->sdf && s->side_stream && !s->serial_contact && s->contact_split) {
        // the contact chain on the side stream next to BOTH gather kernels; only the integration waits for it
        cudaEventRecord(s->ev_fork, st);
        cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0);
        cudaError_t e = enqueue_contact(s, v, s->side_stream, s->contact_first ? s->ev_chain : nullptr);
        cudaEventRecord(s->ev_join, s->side_stream);
        if (s->contact_first) cudaStreamWaitEvent(st, s->ev_chain, 0);
        enqueue_deform(s, v, st);
        enqueue_force(s, v, MODE_EVAL, st);
        cudaStreamWaitEvent(st, s->ev_join, 0);
        k_integrate<<<nblk(s->n, 256), 256, 0, st>>>(v, s->c, mode);
        s->launches++;
        if (e != cudaSuccess && s->enqueue_err == cudaSuccess) s->enqueue_err = e;
        return;
    }
    enqueue_deform_contact(s, v, st);
    enqueue_force(s, v, mode, st);
}

// frame-0 style priming at the current x: elastic force, force_1 and the next position
static int prime(MisSim* s, cudaStream_t st) {
    if (!s->mass_set || !s->material_set) return fail(MIS_E_STATE, "set_mass and set_material must precede startup/step");
    if (!s->p.euler) {
        View v = make_view(s);
        if (s->forces_only) {
            // only the external force / Dirichlet mask changed (sim.py:279-286): the elastic and contact forces of the
            // current frame are still valid; redo force_1 and part_1 (sim.py:247-251) per particle
            k_reintegrate<<<nblk(s->n, 256), 256, 0, st>>>(v, s->c);
            s->launches++;
        } else {
            enqueue_frame(s, v, MODE_PRIME, st);
        }
        enqueue_halo_sync(s, st);
        if (s->enqueue_err != cudaSuccess) { cudaError_t e = s->enqueue_err; s->enqueue_err = cudaSuccess; CK(e); }
        CK_LAUNCH();
    }
    s->dirty = false; s->forces_only = false;
    return MIS_OK;
}

extern "C" int mis_startup(MisSim* s, const float v0[3], void* stream) {
    if (!s || !v0) return fail(MIS_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    s->cur = 0;
    s->v0[0] = v0[0]; s->v0[1] = v0[1]; s->v0[2] = v0[2];
    k_startup<<<nblk(s->n, 256), 256, 0, st>>>(s->x0m, s->n, make_float3(v0[0], v0[1], v0[2]), s->xv[0], s->vel);
    if (s->ref64) ref_startup(*R64(s), s->v0, st);
    CK_LAUNCH(); s->launches++;
    s->started = true; s->dirty = true; s->forces_only = false;
    return MIS_OK;
}

extern "C" int mis_startup_f64(MisSim* s, const double v0[3], void* stream) {
    if (!s || !v0) return fail(MIS_E_INVALID, "null argument");
    if (!s->ref64) return fail(MIS_E_STATE, "mis_startup_f64 needs a scene created with MisParams.fp64 = 1");
    const float vf[3] = {(float)v0[0], (float)v0[1], (float)v0[2]};
    int rc = mis_startup(s, vf, stream);
    if (rc) return rc;
    s->v0[0] = v0[0]; s->v0[1] = v0[1]; s->v0[2] = v0[2];
    ref_startup(*R64(s), s->v0, (cudaStream_t)stream);
    CK_LAUNCH();
    return MIS_OK;
}

extern "C" int mis_set_state(MisSim* s, const float* x_dev, const float* v_dev, void* stream) {
    if (!s || !x_dev || !v_dev) return fail(MIS_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    k_gather_vec3<<<nblk(s->n, 256), 256, 0, st>>>(x_dev, s->perm, s->n, s->xv[s->cur], 1);
    k_gather_vec3<<<nblk(s->n, 256), 256, 0, st>>>(v_dev, s->perm, s->n, s->vel, 0);
    if (s->ref64) {
        ref::kr_gather<double, float><<<nblk(s->n, 256), 256, 0, st>>>(x_dev, s->perm, s->n, 3, R64(s)->s.x);
        ref::kr_gather<double, float><<<nblk(s->n, 256), 256, 0, st>>>(v_dev, s->perm, s->n, 3, R64(s)->s.v);
        R64(s)->primed = false;
    }
    CK_LAUNCH(); s->launches += 2;
    s->started = true; s->dirty = true; s->forces_only = false;
    return MIS_OK;
}

static void enqueue_one_step(MisSim* s, cudaStream_t st) {
    if (s->p.euler) {
        // sim_taichi.py:174-182: forces at frame f, then advance to f+1
        View v = make_view(s);
        enqueue_frame(s, v, MODE_EULER, st);
        s->cur ^= 1;
    } else {
        s->cur ^= 1;                       // part_1 of this step was fused into the previous force kernel
        View v = make_view(s);
        enqueue_frame(s, v, MODE_STEP, st);
        enqueue_halo_sync(s, st);
    }
}

static void enqueue_one_step(MisSim* s, cudaStream_t st);

// launch `steps` steps as one cached CUDA graph (captured on first use for this ping-pong phase)
static int launch_step_graph(MisSim* s, int steps, cudaStream_t st) {
    if (s->graph_stream != st) { drop_graph(s); s->graph_stream = st; }
    MisSim::StepGraph* hit = nullptr;
    for (auto& g : s->graphs) if (g.steps == steps && g.cur == s->cur) hit = &g;
    if (!hit) {
        cudaGraph_t g = nullptr;
        const long long l0 = s->launches;
        const int cur0 = s->cur;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < steps; k++) enqueue_one_step(s, st);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        MisSim::StepGraph sg;
        sg.steps = steps; sg.cur = cur0; sg.launches = s->launches - l0;
        s->cur = cur0; s->launches = l0;
        if (e != cudaSuccess) return fail(MIS_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        e = cudaGraphInstantiate(&sg.exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(MIS_E_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
        s->graphs.push_back(sg);
        hit = &s->graphs.back();
    }
    CK(cudaGraphLaunch(hit->exec, st));
    s->launches += hit->launches;
    if (steps & 1) s->cur ^= 1;                // an odd number of steps flips the ping-pong index
    return MIS_OK;
}

extern "C" int mis_step(MisSim* s, int n_steps, void* stream) {
    if (!s || n_steps < 0) return fail(MIS_E_INVALID, "bad argument");
    if (!s->started) return fail(MIS_E_STATE, "mis_step before mis_startup / mis_set_state");
    cudaStream_t st = (cudaStream_t)stream;
    if (s->ref64) {           // the whole scene steps in double (mis_ref.cuh); 8 launches per step
        if (!s->mass_set || !s->material_set) return fail(MIS_E_STATE, "set_mass and set_material must precede startup/step");
        for (int k = 0; k < n_steps; k++) ref_step(*R64(s), st);
        s->launches += 8LL * n_steps;
        CK_LAUNCH();
        return MIS_OK;
    }
    if (s->dirty) { int rc = prime(s, st); if (rc) return rc; }
    int chunk = s->p.graph_steps > 0 ? s->p.graph_steps : 32;
    int done = 0;
    const bool can_graph = s->p.graph_steps >= 0 && st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
    if (can_graph) {
        while (n_steps - done >= chunk) { int rc = launch_step_graph(s, chunk, st); if (rc) return rc; done += chunk; }
        // the tail (and step(1) loops) go through single-step graphs: one launch instead of 2-40 per step
        for (; done < n_steps; done++) { int rc = launch_step_graph(s, 1, st); if (rc) return rc; }
    }
    for (; done < n_steps; done++) enqueue_one_step(s, st);
    if (s->enqueue_err != cudaSuccess) { cudaError_t e = s->enqueue_err; s->enqueue_err = cudaSuccess; CK(e); }
    CK_LAUNCH();
    return MIS_OK;
}

extern "C" int mis_get_state(MisSim* s, float* x_dev, float* v_dev, void* stream) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    cudaStream_t st = (cudaStream_t)stream;
    if (s->ref64) {           // rounded to fp32; mis_get_f64 returns the doubles
        if (x_dev) ref::kr_scatter<double, float><<<nblk(s->n, 256), 256, 0, st>>>(R64(s)->s.x, s->perm, s->n, 3, x_dev);
        if (v_dev) ref::kr_scatter<double, float><<<nblk(s->n, 256), 256, 0, st>>>(R64(s)->s.v, s->perm, s->n, 3, v_dev);
        CK_LAUNCH(); s->launches += 2;
        return MIS_OK;
    }
    if (x_dev) { k_export_vec3<<<nblk(s->n, 256), 256, 0, st>>>(s->xv[s->cur], s->inv_perm, s->n, x_dev); s->launches++; }
    if (v_dev) { k_export_vec3<<<nblk(s->n, 256), 256, 0, st>>>(s->vel, s->inv_perm, s->n, v_dev); s->launches++; }
    CK_LAUNCH();
    return MIS_OK;
}

extern "C" int mis_get_state_host(MisSim* s, float* x_host, float* v_host, void* stream) {
    NO_FP64("host streaming");
    if (!s) return fail(MIS_E_INVALID, "null sim");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N3 = 3 * (size_t)s->n;
    int rc = mis_get_state(s, x_host ? s->stage : nullptr, v_host ? s->stage + N3 : nullptr, stream);
    if (rc) return rc;
    if (x_host) CK(cudaMemcpyAsync(x_host, s->stage, N3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (v_host) CK(cudaMemcpyAsync(v_host, s->stage + N3, N3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return MIS_OK;
}

// Streaming export: the un-permute kernels run on `stream`, the device->host copies on the library's copy stream, so the
// next step's kernels overlap the transfer.  Two exports may be in flight (double-buffered staging).
extern "C" int mis_get_state_host_async(MisSim* s, float* x_host, float* v_host, void* stream) {
    NO_FP64("host streaming");
    if (!s || (!x_host && !v_host)) return fail(MIS_E_INVALID, "null argument");
    int rc = ensure_copy_stream(s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int k = (int)(s->down_i++ & 1);
    const size_t N3 = 3 * (size_t)s->n;
    CK(cudaStreamWaitEvent(st, s->ev_down[k], 0));                      // the copy that last read this staging slot has finished
    rc = mis_get_state(s, x_host ? s->sstage[k] : nullptr, v_host ? s->sstage[k] + N3 : nullptr, stream);
    if (rc) return rc;
    CK(cudaEventRecord(s->ev_exp[k], st));
    CK(cudaStreamWaitEvent(s->copy_stream, s->ev_exp[k], 0));
    if (x_host) CK(cudaMemcpyAsync(x_host, s->sstage[k], N3 * sizeof(float), cudaMemcpyDeviceToHost, s->copy_stream));
    if (v_host) CK(cudaMemcpyAsync(v_host, s->sstage[k] + N3, N3 * sizeof(float), cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(s->ev_down[k], s->copy_stream));
    return MIS_OK;
}

// Blocks the host until at most `pending_allowed` (0 or 1) of the exports issued by mis_get_state_host_async are still in flight.
extern "C" int mis_wait_state_host(MisSim* s, int pending_allowed) {
    if (!s || pending_allowed < 0 || pending_allowed > 1) return fail(MIS_E_INVALID, "bad argument");
    if (!s->copy_stream || s->down_i == 0) return MIS_OK;
    if (pending_allowed == 0) {
        CK(cudaEventSynchronize(s->ev_down[(s->down_i - 1) & 1]));
        if (s->down_i >= 2) CK(cudaEventSynchronize(s->ev_down[s->down_i & 1]));
    } else if (s->down_i >= 2) {
        CK(cudaEventSynchronize(s->ev_down[s->down_i & 1]));            // slot of export down_i - 2
    }
    return MIS_OK;
}

extern "C" int mis_get_fields(MisSim* s, float* A_dev, float* R_dev, float* F_dev, float* S_dev,
                              float* fel_dev, float* rho_dev, float* vol_dev, void* stream) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    cudaStream_t st = (cudaStream_t)stream;
    if (s->ref64) {
        RefEngine<double>& g = *R64(s);
        ref64_refresh(s, st);
        const int b = nblk(s->n, 256);
        if (A_dev) ref::kr_scatter<double, float><<<b, 256, 0, st>>>(g.s.A, s->perm, s->n, 9, A_dev);
        if (R_dev) ref::kr_scatter<double, float><<<b, 256, 0, st>>>(g.s.R, s->perm, s->n, 9, R_dev);
        if (F_dev) ref::kr_scatter<double, float><<<b, 256, 0, st>>>(g.s.F, s->perm, s->n, 9, F_dev);
        if (S_dev) ref::kr_scatter<double, float><<<b, 256, 0, st>>>(g.s.S, s->perm, s->n, 9, S_dev);
        if (fel_dev) ref::kr_scatter<double, float><<<b, 256, 0, st>>>(g.s.fel, s->perm, s->n, 3, fel_dev);
        if (rho_dev) ref::kr_scatter<double, float><<<b, 256, 0, st>>>(g.s.rho, s->perm, s->n, 1, rho_dev);
        if (vol_dev) ref::kr_scatter<double, float><<<b, 256, 0, st>>>(g.s.vol, s->perm, s->n, 1, vol_dev);
        CK_LAUNCH(); s->launches += 7;
        return MIS_OK;
    }
    if (s->started && s->dirty) { int rc = prime(s, st); if (rc) return rc; }
    View v = make_view(s);
    v.Apq = s->Apq;
    if (A_dev && !s->Apq) return fail(MIS_E_STATE, "A_pq export needs MisParams.keep_fields = 1");
    const int g = nblk(s->n, 256);
    if (R_dev) { k_export_field<<<g, 256, 0, st>>>(v, s->inv_perm, 0, R_dev); s->launches++; }
    if (S_dev) { k_export_field<<<g, 256, 0, st>>>(v, s->inv_perm, 1, S_dev); s->launches++; }
    if (F_dev) { k_export_field<<<g, 256, 0, st>>>(v, s->inv_perm, 2, F_dev); s->launches++; }
    if (A_dev) { k_export_field<<<g, 256, 0, st>>>(v, s->inv_perm, 3, A_dev); s->launches++; }
    if (rho_dev) { k_export_field<<<g, 256, 0, st>>>(v, s->inv_perm, 4, rho_dev); s->launches++; }
    if (vol_dev) { k_export_field<<<g, 256, 0, st>>>(v, s->inv_perm, 5, vol_dev); s->launches++; }
    if (fel_dev) { k_export_vec3<<<g, 256, 0, st>>>(s->fel, s->inv_perm, s->n, fel_dev); s->launches++; }
    CK_LAUNCH();
    return MIS_OK;
}

extern "C" int mis_eval_forces(MisSim* s, const float* x_dev, float* fel_dev, void* stream) {
    if (!s || !x_dev || !fel_dev) return fail(MIS_E_INVALID, "null argument");
    if (!s->mass_set || !s->material_set) return fail(MIS_E_STATE, "set_mass and set_material must precede mis_eval_forces");
    cudaStream_t st = (cudaStream_t)stream;
    if (s->ref64) {
        RefEngine<double>& g = *R64(s);
        ref_statics(g, st);
        ref::kr_gather<double, float><<<nblk(s->n, 256), 256, 0, st>>>(x_dev, s->perm, s->n, 3, g.s.xn);
        ref_eval(g, g.s.xn, g.s.feln, st);
        ref::kr_scatter<double, float><<<nblk(s->n, 256), 256, 0, st>>>(g.s.feln, s->perm, s->n, 3, fel_dev);
        g.primed = false;                     // R, S, F now describe x_dev
        CK_LAUNCH(); s->launches += 5;
        return MIS_OK;
    }
    // scratch position buffer carrying the volumes in .w
    CK(cudaMemcpyAsync(s->scratch4, s->xv[s->cur], (size_t)s->n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
    k_gather_vec3<<<nblk(s->n, 256), 256, 0, st>>>(x_dev, s->perm, s->n, s->scratch4, 1);
    s->launches++;
    View v = make_view(s);
    v.xcur = s->scratch4; v.xnext = s->scratch4 + s->n; v.fel = s->scratch4 + s->n;
    enqueue_deform(s, v, st);
    enqueue_force(s, v, MODE_EVAL, st);
    k_export_vec3<<<nblk(s->n, 256), 256, 0, st>>>(s->scratch4 + s->n, s->inv_perm, s->n, fel_dev);
    s->launches++;
    CK_LAUNCH();
    s->dirty = true; s->forces_only = false;                      // R, S, F now describe x_dev, not the current frame
    return MIS_OK;
}

// positions of frame f+1 (written by the fused part_1): xv[cur ^ 1] once the state is primed
extern "C" int mis_gather_next_positions(MisSim* s, const int* ids_dev, int count, float* x_dev, void* stream) {
    NO_FP64("halo plumbing");
    if (!s || count < 0 || (count > 0 && (!ids_dev || !x_dev))) return fail(MIS_E_INVALID, "bad argument");
    if (!s->started || s->dirty) return fail(MIS_E_STATE, "mis_gather_next_positions needs a primed state (mis_startup + mis_step)");
    if (s->p.euler) return fail(MIS_E_UNSUPPORTED, "halo plumbing supports the velocity-Verlet path only");
    if (count == 0) return MIS_OK;
    k_subset_gather<<<nblk(count, 256), 256, 0, (cudaStream_t)stream>>>(s->xv[s->cur ^ 1], s->inv_perm, ids_dev, count, x_dev);
    CK_LAUNCH(); s->launches++;
    return MIS_OK;
}
extern "C" int mis_scatter_next_positions(MisSim* s, const int* ids_dev, int count, const float* x_dev, void* stream) {
    NO_FP64("halo plumbing");
    if (!s || count < 0 || (count > 0 && (!ids_dev || !x_dev))) return fail(MIS_E_INVALID, "bad argument");
    if (!s->started || s->dirty) return fail(MIS_E_STATE, "mis_scatter_next_positions needs a primed state (mis_startup + mis_step)");
    if (s->p.euler) return fail(MIS_E_UNSUPPORTED, "halo plumbing supports the velocity-Verlet path only");
    if (count == 0) return MIS_OK;
    k_subset_scatter<<<nblk(count, 256), 256, 0, (cudaStream_t)stream>>>(s->xv[s->cur ^ 1], s->inv_perm, ids_dev, count, x_dev);
    CK_LAUNCH(); s->launches++;
    return MIS_OK;
}

extern "C" int mis_set_volumes(MisSim* s, const int* ids_dev, int count, const float* vol_dev, void* stream) {
    NO_FP64("halo plumbing");
    if (!s || count < 0 || (count > 0 && (!ids_dev || !vol_dev))) return fail(MIS_E_INVALID, "bad argument");
    if (!s->mass_set) return fail(MIS_E_STATE, "mis_set_volumes before mis_set_mass");
    cudaStream_t st = (cudaStream_t)stream;
    if (count > 0) {
        k_set_volumes<<<nblk(count, 256), 256, 0, st>>>(s->xv[0], s->xv[1], s->inv_perm, ids_dev, count, vol_dev);
        s->launches++;
    }
    const int blocks = nblk((long long)s->n * s->G, STEP_THREADS);
    if (s->G == 8) k_static_K<8><<<blocks, STEP_THREADS, 0, st>>>(s->x0m, s->xv[0], s->nbr_start, s->nbr, s->n, s->c, s->Ks);
    else if (s->G == 16) k_static_K<16><<<blocks, STEP_THREADS, 0, st>>>(s->x0m, s->xv[0], s->nbr_start, s->nbr, s->n, s->c, s->Ks);
    else k_static_K<32><<<blocks, STEP_THREADS, 0, st>>>(s->x0m, s->xv[0], s->nbr_start, s->nbr, s->n, s->c, s->Ks);
    CK_LAUNCH(); s->launches++;
    s->dirty = true; s->forces_only = false;
    return MIS_OK;
}

extern "C" int mis_export_slots(MisSim* s, const int* ids_dev, int count, int* slots_dev, void* stream) {
    if (!s || count < 0 || (count > 0 && (!ids_dev || !slots_dev))) return fail(MIS_E_INVALID, "bad argument");
    if (count == 0) return MIS_OK;
    k_slots_of<<<nblk(count, 256), 256, 0, (cudaStream_t)stream>>>(s->inv_perm, ids_dev, count, slots_dev);
    CK_LAUNCH(); s->launches++;
    return MIS_OK;
}

// ------------------------------------------------------------------ fused halo push over peer memory
static int ensure_halo_mem(MisSim* s) {
    if (s->halo_mem) return MIS_OK;
    CK(cudaMalloc((void**)&s->halo_mem, 64 * sizeof(unsigned)));
    CK(cudaMemset(s->halo_mem, 0, 64 * sizeof(unsigned)));
    return MIS_OK;
}

extern "C" int mis_halo_local_ptrs(MisSim* s, void** xv0, void** xv1, void** flags) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    int rc = ensure_halo_mem(s);
    if (rc) return rc;
    if (xv0) *xv0 = s->xv[0];
    if (xv1) *xv1 = s->xv[1];
    if (flags) *flags = s->halo_mem;
    return MIS_OK;
}

extern "C" int mis_halo_ipc_handles(MisSim* s, unsigned char* out192) {
    if (!s || !out192) return fail(MIS_E_INVALID, "null argument");
    int rc = ensure_halo_mem(s);
    if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void* ptrs[3] = {s->xv[0], s->xv[1], s->halo_mem};
    for (int k = 0; k < 3; k++) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ptrs[k]));
        memcpy(out192 + 64 * k, &h, 64);
    }
    return MIS_OK;
}

extern "C" int mis_ipc_open(const unsigned char* handle64, void** dev_ptr) {
    if (!handle64 || !dev_ptr) return fail(MIS_E_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MIS_OK;
}

extern "C" int mis_ipc_close(void* dev_ptr) {
    if (dev_ptr) CK(cudaIpcCloseMemHandle(dev_ptr));
    return MIS_OK;
}

extern "C" int mis_halo_connect(MisSim* s, int n_peers, void* const* peer_xv0, void* const* peer_xv1, void* const* peer_flag,
                                int n_push, const int* push_ids_dev, const int* push_peer_dev, const int* push_slot_dev,
                                int n_ghost, const int* ghost_ids_dev, const int* ghost_layer_dev, void* stream) {
    NO_FP64("halo plumbing");
    if (!s || n_peers < 0 || n_peers > MIS_MAX_PEERS || n_push < 0 || n_ghost < 0) return fail(MIS_E_INVALID, "mis_halo_connect: bad argument");
    if (n_peers > 0 && (!peer_xv0 || !peer_xv1 || !peer_flag)) return fail(MIS_E_INVALID, "mis_halo_connect: null peer table");
    if (n_push > 0 && (!push_ids_dev || !push_peer_dev || !push_slot_dev)) return fail(MIS_E_INVALID, "mis_halo_connect: null push list");
    if (n_ghost > 0 && !ghost_ids_dev) return fail(MIS_E_INVALID, "mis_halo_connect: null ghost list");
    if (s->p.euler) return fail(MIS_E_UNSUPPORTED, "halo plumbing supports the velocity-Verlet path only");
    if ((long long)s->n >= (1ll << PUSH_SLOT_BITS)) return fail(MIS_E_UNSUPPORTED, "halo push supports up to 2^28 particles per rank");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_halo_mem(s);
    if (rc) return rc;
    if (!s->push) CK(dalloc(&s->push, (size_t)s->n));
    CK(cudaMemsetAsync(s->push, 0xFF, (size_t)s->n * sizeof(int2), st));
    const int m = n_push > n_ghost ? n_push : n_ghost;
    if (m > 0) {
        k_push_fill<<<nblk(m, 256), 256, 0, st>>>(s->push, s->inv_perm, n_push, push_ids_dev, push_peer_dev, push_slot_dev, n_ghost, ghost_ids_dev, ghost_layer_dev);
        CK_LAUNCH(); s->launches++;
    }
    for (int p = 0; p < MIS_MAX_PEERS; p++) {
        s->peer_xv[0][p] = p < n_peers ? (float4*)peer_xv0[p] : nullptr;
        s->peer_xv[1][p] = p < n_peers ? (float4*)peer_xv1[p] : nullptr;
        s->peer_flag[p] = p < n_peers ? (unsigned*)peer_flag[p] : nullptr;
    }
    s->halo_peers = n_peers;
    s->halo_on = true;
    const char* e = getenv("MIS_HALO_TIMEOUT_MS");
    if (e && atof(e) > 0.0) s->halo_timeout_ns = (unsigned long long)(atof(e) * 1e6);
    drop_graph(s);
    s->dirty = true; s->forces_only = false;
    CK(cudaStreamSynchronize(st));
    return MIS_OK;
}

extern "C" int mis_halo_disconnect(MisSim* s) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    drop_graph(s);
    s->halo_on = false; s->halo_peers = 0;
    s->dirty = true; s->forces_only = false;
    return MIS_OK;
}

extern "C" int mis_halo_set_wait(MisSim* s, int wait) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    drop_graph(s);
    s->halo_wait = wait != 0;
    return MIS_OK;
}

extern "C" int mis_halo_status(MisSim* s, void* stream, int* err, long long* exchanges) {
    if (!s) return fail(MIS_E_INVALID, "null sim");
    if (err) *err = 0;
    if (exchanges) *exchanges = s->exchanges;
    if (!s->halo_mem) return MIS_OK;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    unsigned host[2] = {0u, 0u};                                     // epoch, error flag
    CK(cudaMemcpy(host, s->halo_mem + MIS_MAX_PEERS, sizeof host, cudaMemcpyDeviceToHost));
    if (err) *err = (int)host[1];
    if (exchanges) *exchanges = (long long)host[0];                  // counted on the device: graph replays included
    return MIS_OK;
}

extern "C" int mis_accumulate_loss(MisSim* s, const float* target_x_dev, const float* target_v_dev, double* loss_dev, void* stream) {
    NO_FP64("mis_accumulate_loss (use mis_rollout_grad)");
    if (!s || !target_x_dev || !target_v_dev || !loss_dev) return fail(MIS_E_INVALID, "null argument");
    if (!s->started) return fail(MIS_E_STATE, "mis_accumulate_loss before mis_startup / mis_set_state");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = 296;                                        // 2 per SM
    if (!s->loss_partial) CK(dalloc(&s->loss_partial, (size_t)blocks));
    double* partial = s->loss_partial;
    k_loss_partial<<<blocks, 256, 0, st>>>(s->xv[s->cur], s->vel, s->inv_perm, target_x_dev, target_v_dev, s->n, s->p.dt, partial);
    k_loss_final<<<1, 1, 0, st>>>(partial, blocks, loss_dev);
    CK_LAUNCH(); s->launches += 2;
    return MIS_OK;
}

extern "C" int mis_set_gather_mode(MisSim* s, int mode, void* stream) {
    if (!s || mode < 0 || mode > 2) return fail(MIS_E_INVALID, "mis_set_gather_mode: mode must be 0, 1 or 2");
    drop_graph(s);
    s->tile.want_d = mode >= 1; s->tile.want_f = mode == 1;
    s->dirty = true; s->forces_only = false;
    if (mode && s->built && !s->tile.ok) { int rc = build_tiles(s, (cudaStream_t)stream); if (rc) return rc; }
    if (mode && !s->tile.ok) {
        s->tile.want_d = s->tile.want_f = false;
        return fail(MIS_E_UNSUPPORTED, "a 27-cell neighbourhood of this scene exceeds the largest shared-memory tile");
    }
    return MIS_OK;
}

extern "C" int mis_get_gather_info(MisSim* s, int out[8]) {
    if (!s || !out) return fail(MIS_E_INVALID, "null argument");
    out[0] = use_tiles_d(s) ? (use_tiles_f(s) ? 1 : 2) : 0; out[1] = s->tile.n_active; out[2] = s->tile.max_tile; out[3] = s->tile.max_own;
    out[4] = s->tile.cap_d; out[5] = s->tile.cap_f; out[6] = (int)(s->tile.total_blocks >> 31 ? 0x7fffffff : s->tile.total_blocks); out[7] = TILE_BLOCK;
    return MIS_OK;
}

extern "C" long long mis_launch_count(MisSim* s) { return s ? s->launches : 0; }

// n_steps steps launched one kernel at a time with a CUDA event pair around each launch on
// `stream`; returns the summed device time per kernel family.  Synchronises the host.
extern "C" int mis_profile_step(MisSim* s, int n_steps, void* stream, double* ms_deform, double* ms_force) {
    NO_FP64("mis_profile_step");
    if (!s || n_steps <= 0) return fail(MIS_E_INVALID, "bad argument");
    if (!s->started) return fail(MIS_E_STATE, "mis_profile_step before mis_startup / mis_set_state");
    if (s->sdf) return fail(MIS_E_UNSUPPORTED, "mis_profile_step times the two gather kernels alone: remove the obstacle first (mis_set_sdf_contact(sim, NULL, ...))");
    cudaStream_t st = (cudaStream_t)stream;
    if (s->dirty) { int rc = prime(s, st); if (rc) return rc; }
    std::vector<cudaEvent_t> ev(3 * (size_t)n_steps);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    for (int k = 0; k < n_steps; k++) {
        if (!s->p.euler) s->cur ^= 1;
        View v = make_view(s);
        CK(cudaEventRecord(ev[3 * k + 0], st));
        enqueue_deform(s, v, st);
        CK(cudaEventRecord(ev[3 * k + 1], st));
        enqueue_force(s, v, s->p.euler ? MODE_EULER : MODE_STEP, st);
        CK(cudaEventRecord(ev[3 * k + 2], st));
        if (s->p.euler) s->cur ^= 1; else enqueue_halo_sync(s, st);
    }
    CK(cudaStreamSynchronize(st));
    double a = 0, b = 0;
    for (int k = 0; k < n_steps; k++) {
        float t0 = 0, t1 = 0;
        cudaEventElapsedTime(&t0, ev[3 * k + 0], ev[3 * k + 1]);
        cudaEventElapsedTime(&t1, ev[3 * k + 1], ev[3 * k + 2]);
        a += t0; b += t1;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (ms_deform) *ms_deform = a;
    if (ms_force) *ms_force = b;
    CK_LAUNCH();
    return MIS_OK;
}

// ------------------------------------------------------------------ DeepSDF (deepsdf.py:9-41)
extern "C" int mis_sdf_create(int n_layers, const int* dims, const float* const* g_dev, const float* const* v_dev,
                              const float* const* bias_dev, void* stream, MisSdf** out) {
    if (!out || !dims || !g_dev || !v_dev || !bias_dev || n_layers < 3) return fail(MIS_E_INVALID, "mis_sdf_create: bad argument");
    const int H = dims[1];
    if (dims[0] != 3 || dims[n_layers] != 1) return fail(MIS_E_UNSUPPORTED, "mis_sdf_create: network must map 3 -> 1 (deepsdf.py:13,37)");
    for (int l = 1; l < n_layers; l++) if (dims[l] != H) return fail(MIS_E_UNSUPPORTED, "mis_sdf_create: hidden widths must be equal");
    if (H % SDF_BN != 0) return fail(MIS_E_UNSUPPORTED, "mis_sdf_create: hidden width must be a multiple of 256");
    cudaStream_t st = (cudaStream_t)stream;
    MisSdf* s = new MisSdf();
    s->L = n_layers; s->H = H;
    { int dev = 0, sms = 0; if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) s->num_sms = sms; }
#define SALLOC(ptr, cnt) do { cudaError_t e_ = cudaMalloc((void**)&(ptr), (size_t)(cnt) * sizeof(float)); if (e_ != cudaSuccess) { int r_ = fail(MIS_E_CUDA, std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_)); sdf_free(s); return r_; } } while (0)
    SALLOC(s->W0, (size_t)H * 3); SALLOC(s->b0, H); SALLOC(s->wl, H); SALLOC(s->bl, 1);
    k_sdf_pack_weights<<<H, 256, 0, st>>>(g_dev[0], v_dev[0], H, 3, s->W0, nullptr, nullptr);
    cudaMemcpyAsync(s->b0, bias_dev[0], H * sizeof(float), cudaMemcpyDeviceToDevice, st);
    const size_t hidden = (size_t)(n_layers - 2);
    SALLOC(s->Wslab, hidden * 2 * (size_t)H * H);
    {   // per-step contact query: the weights (hidden * 8 MB at H = 1024) are re-read every step while the gather kernels stream
        // > 100 MB of neighbour lists through L2 in between; keep them in the persisting part of L2 (MIS_SDF_L2_PIN=0 disables)
        const char* e = getenv("MIS_SDF_L2_PIN");
        int dev = 0, max_persist = 0, max_win = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        const size_t need = hidden * 2 * (size_t)H * H * sizeof(float);
        if (!(e && e[0] == '0') && max_persist > 0 && (size_t)max_win >= 2 * (size_t)H * H * sizeof(float)) {
            size_t cur_lim = 0;
            cudaDeviceGetLimit(&cur_lim, cudaLimitPersistingL2CacheSize);
            const size_t want = need < (size_t)max_persist ? need : (size_t)max_persist;
            if (cur_lim < want) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            s->l2_pin = true;
        }
    }
    SALLOC(s->bslab, hidden * (size_t)H);
    { const char* e = getenv("MIS_SDF_CHAIN"); if (e && e[0] == '0') s->chain_mode = 0; }
    for (int l = 1; l < n_layers - 1; l++) {
        float *hi = s->Wslab + (size_t)(l - 1) * 2 * H * H, *lo = hi + (size_t)H * H, *b = s->bslab + (size_t)(l - 1) * H;
        s->Whi.push_back(hi); s->Wlo.push_back(lo);
        s->bh.push_back(b);
        k_sdf_pack_weights<<<H, 256, 0, st>>>(g_dev[l], v_dev[l], H, H, nullptr, hi, lo);
        cudaMemcpyAsync(b, bias_dev[l], H * sizeof(float), cudaMemcpyDeviceToDevice, st);
    }
    k_sdf_pack_weights<<<1, 256, 0, st>>>(g_dev[n_layers - 1], v_dev[n_layers - 1], 1, H, s->wl, nullptr, nullptr);
    cudaMemcpyAsync(s->bl, bias_dev[n_layers - 1], sizeof(float), cudaMemcpyDeviceToDevice, st);
#undef SALLOC
    s->launches += n_layers;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { int r = fail(MIS_E_CUDA, std::string("mis_sdf_create: ") + cudaGetErrorString(e)); sdf_free(s); return r; }
    *out = s;
    return MIS_OK;
}

extern "C" int mis_sdf_destroy(MisSdf* s) { sdf_free(s); return MIS_OK; }

static SdfXform make_xform(const float* xform_host) {
    SdfXform xf;
    for (int k = 0; k < 9; k++) xf.A[k] = xform_host ? xform_host[k] : ((k % 4 == 0) ? 1.f : 0.f);
    for (int k = 0; k < 3; k++) xf.t[k] = xform_host ? xform_host[9 + k] : 0.f;
    return xf;
}

extern "C" int mis_sdf_query(MisSdf* s, const float* points_dev, int n, const float* xform_host,
                             float* sdf_dev, float* grad_dev, float fd_eps, void* stream) {
    if (!s || !points_dev || n < 0 || (!sdf_dev && !grad_dev)) return fail(MIS_E_INVALID, "mis_sdf_query: bad argument");
    if (grad_dev && !(fd_eps > 0.f)) return fail(MIS_E_INVALID, "mis_sdf_query: fd_eps must be positive when a gradient is requested");
    if (n == 0) return MIS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const SdfXform xf = make_xform(xform_host);
    const int chunk = 1 << 16;                     // rows per pass: 4 activation planes of chunk x H floats stay resident
    CK(sdf_reserve(s, n < chunk ? n : chunk));
    for (int r0 = 0; r0 < n; r0 += chunk) {
        const int rows = n - r0 < chunk ? n - r0 : chunk;
        const float* pts = points_dev + 3 * (size_t)r0;
        if (!grad_dev) {
            CK(sdf_forward(s, pts, nullptr, rows, nullptr, xf, make_float3(0.f, 0.f, 0.f), sdf_dev + r0, st));
        } else {
            const float3 shifts[4] = {make_float3(0.f, 0.f, 0.f), make_float3(fd_eps, 0.f, 0.f), make_float3(0.f, fd_eps, 0.f), make_float3(0.f, 0.f, fd_eps)};
            for (int q = 0; q < 4; q++) CK(sdf_forward(s, pts, nullptr, rows, nullptr, xf, shifts[q], s->vals + (size_t)q * s->cap, st));
            k_sdf_fd_grad<<<nblk(rows, 256), 256, 0, st>>>(s->vals, s->cap, rows, 1.f / fd_eps, xf, sdf_dev ? sdf_dev + r0 : nullptr, grad_dev + 3 * (size_t)r0);
            s->launches++;
        }
    }
    CK_LAUNCH();
    return MIS_OK;
}

extern "C" int mis_sdf_set_gemm_path(MisSdf* s, int path) {
    if (!s || path < 0 || path > 3) return fail(MIS_E_INVALID, "mis_sdf_set_gemm_path: path must be 0, 1, 2 or 3");
    if (path == 3 && (s->H % CH_BN != 0 || s->H / CH_BN > 15)) return fail(MIS_E_UNSUPPORTED, "chain kernel supports hidden widths up to 1024");
    if ((path == 1 || path == 3) && s->H > SK_MAX_KBS * SK_SPLIT * SDF_BK) return fail(MIS_E_UNSUPPORTED, "split-K kernel supports hidden widths up to 1024");
    s->force_path = path;
    return MIS_OK;
}

extern "C" long long mis_sdf_launch_count(MisSdf* s, long long* gemm_launches) {
    if (!s) return 0;
    if (gemm_launches) *gemm_launches = s->gemm_launches;
    return s->launches;
}

extern "C" int mis_sdf_profile_gemm(MisSdf* s, int m, int reps, void* stream, double* ms_total) {
    if (!s || m <= 0 || reps <= 0 || !ms_total) return fail(MIS_E_INVALID, "mis_sdf_profile_gemm: bad argument");
    if (s->Whi.empty()) return fail(MIS_E_STATE, "network has no hidden layer");
    cudaStream_t st = (cudaStream_t)stream;
    CK(sdf_reserve(s, m));
    CK(cudaFuncSetAttribute(k_sdf_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_sdf_gemm_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES));
    const int m_pad = (m + 127) / 128 * 128;
    CK(cudaMemsetAsync(s->act[0][0], 0, (size_t)m_pad * s->H * sizeof(float), st));
    CK(cudaMemsetAsync(s->act[0][1], 0, (size_t)m_pad * s->H * sizeof(float), st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const bool skinny = s->force_path ? (s->force_path == 1) : (s->H <= SK_MAX_KBS * SK_SPLIT * SDF_BK && m <= 1024);
    const int tiles = (s->H / SDF_BN) * (m_pad / SDF_BM);
    const int grid = tiles < s->num_sms ? tiles : s->num_sms;
    auto one = [&]() {
        if (skinny) k_sdf_gemm_sk<<<(s->H / SK_BN) * SK_SPLIT, SDF_THREADS, SK_SMEM_BYTES, st>>>(s->act[0][0], s->act[0][1], s->Whi[0], s->Wlo[0], s->bh[0], s->H, s->H, s->act[1][0], s->act[1][1], m, nullptr);
        else k_sdf_gemm<<<grid, SDF_THREADS, SDF_SMEM_BYTES, st>>>(s->act[0][0], s->act[0][1], s->Whi[0], s->Wlo[0], s->bh[0], s->H, s->H, s->act[1][0], s->act[1][1], m_pad, nullptr);
    };
    one();
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; r++) one();
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    s->launches += reps + 1; s->gemm_launches += reps + 1;
    *ms_total = ms;
    CK_LAUNCH();
    return MIS_OK;
}

extern "C" int mis_set_sdf_contact(MisSim* s, MisSdf* sdf, const float* xform_host, const float* bbox_host, float fd_eps, void* stream) {
    NO_FP64("obstacle contact");
    if (!s) return fail(MIS_E_INVALID, "null sim");
    drop_graph(s);
    s->dirty = true; s->forces_only = false;
    if (!sdf) { s->sdf = nullptr; return MIS_OK; }
    if (!bbox_host || !(fd_eps > 0.f)) return fail(MIS_E_INVALID, "mis_set_sdf_contact: bbox and a positive fd_eps are required");
    const size_t N = (size_t)s->n;
    if (!s->fcon) {
        CK(dalloc(&s->con_idx, N)); CK(dalloc(&s->con_idx2, N)); CK(dalloc(&s->con_count, (size_t)4));
        CK(dalloc(&s->con_pts, 3 * N)); CK(dalloc(&s->con_pts2, 3 * N)); CK(dalloc(&s->con_s0, N)); CK(dalloc(&s->fcon, N));
    }
    if (!s->side_stream) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));                  // hi = numerically lowest = highest priority
        CK(cudaStreamCreateWithPriority(&s->side_stream, cudaStreamNonBlocking, hi));
        CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->ev_chain, cudaEventDisableTiming));
        const char* e = getenv("MIS_SERIAL_CONTACT");
        s->serial_contact = e && e[0] == '1';
        const char* f = getenv("MIS_CONTACT_FIRST");
        s->contact_first = !(f && f[0] == '0');
    }
    {   // The chain (~0.2 ms, latency-bound) hides behind the deformation kernel alone once that runs longer than the chain
        // (n > ~3e5); below that the split buys the overlap with the force gather for one more pass over ~150 B per particle.
        const int sp = env_int("MIS_CONTACT_SPLIT", -1);
        s->contact_split = sp < 0 ? s->n <= 400000 : sp != 0;
    }
    CK(cudaMemsetAsync(s->con_count, 0, 4 * sizeof(int), (cudaStream_t)stream));
    {   // rows one pass of the chain can take: the activations are 16 KB per row (4 buffers x H floats), so the capacity is bounded
        // (not n: 10 M particles would ask for 164 GB); an over-full bounding box or contact band raises the overflow flag
        const int want = env_int("MIS_CONTACT_ROWS", 32768);
        int cap = want < 384 ? 384 : want;
        if (cap > s->n) cap = s->n;
        s->con_cap = (cap + 127) / 128 * 128;
    }
    CK(sdf_reserve(sdf, s->con_cap));
    CK(cudaFuncSetAttribute(k_sdf_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, SDF_SMEM_BYTES));
    s->sdf = sdf; s->sdf_xf = make_xform(xform_host); s->sdf_eps = fd_eps;
    s->sdf_lo = make_float3(bbox_host[0], bbox_host[1], bbox_host[2]);
    s->sdf_hi = make_float3(bbox_host[3], bbox_host[4], bbox_host[5]);
    return MIS_OK;
}

extern "C" int mis_get_contact_force(MisSim* s, float* f_dev, void* stream) {
    if (!s || !f_dev) return fail(MIS_E_INVALID, "null argument");
    if (!s->sdf) return fail(MIS_E_STATE, "no obstacle set (mis_set_sdf_contact)");
    cudaStream_t st = (cudaStream_t)stream;
    if (s->started && s->dirty) { int rc = prime(s, st); if (rc) return rc; }
    k_export_vec3<<<nblk(s->n, 256), 256, 0, st>>>(s->fcon, s->inv_perm, s->n, f_dev);
    CK_LAUNCH(); s->launches++;
    return MIS_OK;
}

extern "C" int mis_get_contact_count(MisSim* s, void* stream, int* count) {
    if (!s || !count) return fail(MIS_E_INVALID, "null argument");
    count[0] = count[1] = 0;
    if (!s->sdf) return MIS_OK;
    int host[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(host, s->con_count, 4 * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    count[0] = host[0]; count[1] = host[1];
    if (host[3]) {
        CK(cudaMemsetAsync(s->con_count + 3, 0, sizeof(int), (cudaStream_t)stream));
        return fail(MIS_E_UNSUPPORTED, "obstacle contact overflow: more particles in the obstacle's bounding box (or 3 x contact band) than the "
                                       "chain's row capacity (MIS_CONTACT_ROWS, " + std::to_string(s->con_cap) + "); their contact force was dropped");
    }
    return MIS_OK;
}
