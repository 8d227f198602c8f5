// frx_capi.cu -- the C ABI (include/frx.h) over the sm_100a kernels.  Host C++ only; no torch types.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>       // header-only; a no-op unless a tool (ncu --nvtx, Nsight Systems) is attached

#include "frx.h"
#include "frx_device.cuh"

// NVTX range over a phase of a C-ABI call: set-up, enqueue (copies + launches), wait (stream sync + result record)
struct FrxRange {
    explicit FrxRange(const char* name) { nvtxRangePushA(name); }
    ~FrxRange() { nvtxRangePop(); }
    FrxRange(const FrxRange&) = delete;
    FrxRange& operator=(const FrxRange&) = delete;
};

// launchers implemented in frx_kernels.cu
size_t frx_eval_smem_bytes(int Mpad, int Nt);
int frx_memo_pitch_host(int Nt);
cudaError_t frx_launch_eval(const FrxKernelArgs& a, int Nt, int grid, cudaStream_t st);
cudaError_t frx_eval_occupancy(int Mpad, int Nt, int* blocks_per_sm);
cudaError_t frx_launch_eval_batched(const FrxKernelArgs* h_agents, const FrxKernelArgs* d_agents, const int* d_cta_begin,
                                    int n_agents, int max_Mpad, int Nt, int grid, cudaStream_t st);
void frx_launch_obstacle_prep(int O, int T, int Tp, const double* pos, const double* cov, const double* theta,
                              const double* hl, const double* hw, double* obs, cudaStream_t st);
void frx_launch_static_prep(int B, const double* obb, double* out, cudaStream_t st);
void frx_launch_obstacle_compact(int O, int Tp, int Nt, const double* obs, const int* obs_len, double ox, double oy, double* pred,
                                 double* hull, float4* hull32, int* n_pred, int* n_hull, cudaStream_t st);
void frx_launch_static_cull(int B, const double* sobb, double ox, double oy, float4* out, cudaStream_t st);
void frx_launch_prob_records(int O, int T, int Tp, const double* pos, const double* cov, const double* theta, const double* hl,
                             const int* obs_len, double* out, cudaStream_t st);
double frx_measure_fp64_peak(int sm_count, cudaStream_t st, cudaError_t* err);
void frx_launch_reference_tables(int M, int Mpad, const double* xy, double* tab, double* scratch, cudaStream_t st);
void frx_launch_initial_state(int M, int Mpad, const double* tab, const double* in, double wheelbase, int low, double* out,
                              cudaStream_t st);
void frx_launch_collision_counter(long long N, long long row_base, const double* total, const uint32_t* flags,
                                  const FrxBest* winner, unsigned long long* counters, int grid, cudaStream_t st);
void frx_launch_gather(const double* states, long long Np, int Nt, int Ntp, const long long* idx, long long first,
                       long long n_idx, uint32_t mask, double* out, cudaStream_t st);
void frx_features(const FrxKernelArgs& a, bool* obs, bool* xcost);
int frx_pick_seg(long long n_rows, int sm_count);
cudaError_t frx_launch_obstacle_pass(FrxKernelArgs& a, int sm_count, cudaStream_t st, int* launches);
size_t frx_obstacle_scratch_elems(long long N);
int frx_obstacle_pass_max_grid(int sm_count);

void frx_launch_selftest_fdiv(long long n, const double* a, const double* b, double* q1, double* q2, cudaStream_t st);
void frx_launch_selftest_divc(long long n, const double* a, double b, double* q1, double* q2, cudaStream_t st);

namespace {

typedef FrxHostResult HostResult;   // written by the eval kernel's last CTA into mapped host memory

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct frx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr, evkm = nullptr;
    bool split_last = false;           // the last plan ran the obstacle pass as its own kernel (evkm lies between the two)
    bool counted_last = false;         // the last plan ran frx_collision_counter_kernel (else its result record still holds
                                       // the count the PREVIOUS plan left in the device counter: report 0)
    std::string err;
    int sm_count = 148;
    int max_smem_optin = 0;

    frx_params prm{};
    bool have_params = false, have_ref = false, have_tables = false;
    double kappa_max = 0.0;

    DevBuf<double> ref; int M = 0, Mpad = 0; double inv_step = 0.0;
    DevBuf<double> Ttab; DevBuf<int> Tlen; DevBuf<double> tpow; int nT = 0, tpitch = 0;
    DevBuf<double> oprob; bool prob_ready = false;
    DevBuf<double> opred, ohull; DevBuf<float4> ohull32, sobb32; DevBuf<int> on_pred, on_hull; int compact_Nt = 0;
    double origin_x = 0.0, origin_y = 0.0;     // frame of the fp32 cull records: first vertex of the reference path
    bool sobb32_dirty = true;
    DevBuf<double> obs, raw_pos, raw_cov, raw_theta, raw_hl, raw_hw; DevBuf<int> obs_len; int O = 0, T = 0, Tp = 0;
    DevBuf<double> obs_pos; int n_obs_pos = 0;
    DevBuf<double> sobb, raw_sobb; int B = 0;
    DevBuf<double> sampling, grid;
    double* grid_stage = nullptr; size_t grid_stage_cap = 0;    // pinned staging of the three grid axes
    DevBuf<double> states, costs, total; DevBuf<uint32_t> flags; DevBuf<int> traj_len;
    DevBuf<unsigned long long> blockcnt;
    DevBuf<double> obs_part; DevBuf<uint32_t> obs_hit;     // scratch of the step-chunked obstacle pass
    DevBuf<FrxBest> blockbest, winner; DevBuf<unsigned long long> counters;
    DevBuf<long long> gidx; DevBuf<double> gout;
    DevBuf<FrxKernelArgs> batch_args; DevBuf<int> batch_cta;
    HostResult* h_res = nullptr;       // pinned + mapped
    HostResult* d_res = nullptr;       // device address of h_res
    FrxXchgSlot* xchg_host = nullptr;  // the node's shared exchange page (host view / device view), see frx_set_exchange
    FrxXchgSlot* xchg_dev = nullptr;
    int xchg_rank = 0, xchg_world = 1;
    bool xchg_registered = false;
    unsigned long long xchg_epoch = 0; // epoch of the last plan enqueued with the exchange attached
    bool pending = false;              // an asynchronous plan is in flight on `stream`
    bool counters_dirty = true;        // counters must be zeroed before the next launch (first use / after an error)

    long long lastN = 0, lastNp = 0; int lastK = 0, lastNt = 0, lastNtp = 0;
    bool last_all_fields = false;     // the last plan materialised all 14 state planes
    int last_launches = 0;            // kernels the last plan launched (eval, obstacle, collision counter, set-up)
    int occ_Mpad = -1, occ_Nt = -1, occ_blocks = 1;
};

#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);            \
            return (e__ == cudaErrorMemoryAllocation) ? FRX_ERR_NOMEM : FRX_ERR_CUDA;  \
        }                                                                              \
    } while (0)

#define REQUIRE(cond, msg)        \
    do {                          \
        if (!(cond)) {            \
            ctx->err = (msg);     \
            return FRX_ERR_INVALID; \
        }                         \
    } while (0)

static inline int nchunk_for(int Nt) { return (Nt + 31) / 32; }
static inline int pitch_for(int Nt) { return (Nt + 3) & ~3; }

extern "C" {

int frx_abi_version(void) { return FRX_ABI_VERSION; }

int frx_create(int device_ordinal, frx_ctx** out) {
    if (!out) return FRX_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_ordinal < 0 || device_ordinal >= n)
        return FRX_ERR_CUDA;   // no CUDA device: fail loudly, there is no CPU fallback
    frx_ctx* ctx = new (std::nothrow) frx_ctx();
    if (!ctx) return FRX_ERR_NOMEM;
    ctx->device = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) { delete ctx; return FRX_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { delete ctx; return FRX_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FRX_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->evk0); cudaEventCreate(&ctx->evk1); cudaEventCreate(&ctx->evkm);
    if (cudaHostAlloc((void**)&ctx->h_res, sizeof(HostResult), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->d_res, ctx->h_res, 0) != cudaSuccess) {
        frx_destroy(ctx);
        return FRX_ERR_NOMEM;
    }
    memset(ctx->h_res, 0, sizeof(HostResult));
    if (ctx->winner.reserve(1) != cudaSuccess || ctx->counters.reserve(FRX_NUM_COUNTERS) != cudaSuccess) {
        frx_destroy(ctx);
        return FRX_ERR_NOMEM;
    }
    *out = ctx;
    return FRX_OK;
}

int frx_destroy(frx_ctx* ctx) {
    if (!ctx) return FRX_OK;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    ctx->ref.release(); ctx->Ttab.release(); ctx->Tlen.release(); ctx->tpow.release();
    ctx->oprob.release(); ctx->opred.release(); ctx->ohull.release(); ctx->ohull32.release(); ctx->sobb32.release(); ctx->on_pred.release(); ctx->on_hull.release();
    ctx->obs.release(); ctx->raw_pos.release(); ctx->raw_cov.release(); ctx->raw_theta.release();
    ctx->raw_hl.release(); ctx->raw_hw.release(); ctx->obs_len.release(); ctx->obs_pos.release();
    ctx->sobb.release(); ctx->raw_sobb.release(); ctx->sampling.release(); ctx->grid.release();
    ctx->states.release(); ctx->costs.release(); ctx->total.release(); ctx->flags.release(); ctx->traj_len.release();
    ctx->blockcnt.release(); ctx->obs_part.release(); ctx->obs_hit.release(); ctx->blockbest.release(); ctx->winner.release(); ctx->counters.release(); ctx->gidx.release(); ctx->gout.release(); ctx->batch_args.release(); ctx->batch_cta.release();
    if (ctx->xchg_registered) cudaHostUnregister(ctx->xchg_host);
    if (ctx->grid_stage) cudaFreeHost(ctx->grid_stage);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->evk0) cudaEventDestroy(ctx->evk0);
    if (ctx->evk1) cudaEventDestroy(ctx->evk1);
    if (ctx->evkm) cudaEventDestroy(ctx->evkm);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return FRX_OK;
}

const char* frx_last_error(const frx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int frx_set_stream(frx_ctx* ctx, void* cuda_stream) {
    if (!ctx) return FRX_ERR_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return FRX_OK;
}

int frx_synchronize(frx_ctx* ctx) {
    if (!ctx) return FRX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return FRX_OK;
}

int frx_set_reference(frx_ctx* ctx, int32_t M, const double* ref_pos, const double* ref_theta,
                      const double* ref_curv, const double* ref_curv_d, const double* ref_x, const double* ref_y) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(M >= 2 && ref_pos && ref_theta && ref_curv && ref_curv_d && ref_x && ref_y, "frx_set_reference: bad arguments");
    CK(cudaSetDevice(ctx->device));
    int Mpad = (M + 1) & ~1;
    REQUIRE(frx_eval_smem_bytes(Mpad, ctx->have_params ? ctx->prm.N + 1 : 64) <= (size_t)ctx->max_smem_optin,
            "frx_set_reference: reference path too long for the shared-memory table");
    std::vector<double> h((size_t)6 * Mpad, 0.0);
    const double* src[6] = {ref_pos, ref_theta, ref_curv, ref_curv_d, ref_x, ref_y};
    for (int k = 0; k < 6; ++k) memcpy(h.data() + (size_t)k * Mpad, src[k], sizeof(double) * M);
    CK(ctx->ref.reserve(h.size()));
    // on the context's stream, then synchronised: ordered against earlier plans AND complete before the (pageable)
    // staging vector dies -- the legacy stream does not order against a non-blocking stream
    CK(cudaMemcpyAsync(ctx->ref.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->M = M; ctx->Mpad = Mpad; ctx->have_ref = true;
    ctx->origin_x = ref_x[0]; ctx->origin_y = ref_y[0];
    ctx->compact_Nt = 0; ctx->sobb32_dirty = true;     // the fp32 cull records live in the frame of the reference path
    ctx->inv_step = (double)(M - 1) / (ref_pos[M - 1] - ref_pos[0]);
    return FRX_OK;
}

int frx_set_reference_polyline(frx_ctx* ctx, int32_t M, const double* xy) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(M >= 3 && xy, "frx_set_reference_polyline: need a polyline of at least 3 vertices");
    CK(cudaSetDevice(ctx->device));
    const int Mpad = (M + 1) & ~1;
    REQUIRE(frx_eval_smem_bytes(Mpad, ctx->have_params ? ctx->prm.N + 1 : 64) <= (size_t)ctx->max_smem_optin,
            "frx_set_reference_polyline: reference path too long for the shared-memory table");
    DevBuf<double> in, scratch;
    CK(in.reserve((size_t)2 * M)); CK(scratch.reserve((size_t)4 * M)); CK(ctx->ref.reserve((size_t)6 * Mpad));
    CK(cudaMemcpyAsync(in.p, xy, sizeof(double) * 2 * M, cudaMemcpyHostToDevice, ctx->stream));
    frx_launch_reference_tables(M, Mpad, in.p, ctx->ref.p, scratch.p, ctx->stream);
    CK(cudaGetLastError());
    double ends[2] = {0.0, 0.0};      // ref_pos[0] and ref_pos[M-1] for the segment-search guess
    CK(cudaMemcpyAsync(&ends[1], ctx->ref.p + (M - 1), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    in.release(); scratch.release();
    ctx->M = M; ctx->Mpad = Mpad; ctx->have_ref = true;
    ctx->inv_step = (double)(M - 1) / (ends[1] - ends[0]);
    ctx->origin_x = xy[0]; ctx->origin_y = xy[1];
    ctx->compact_Nt = 0; ctx->sobb32_dirty = true;
    return FRX_OK;
}

int frx_get_reference(frx_ctx* ctx, int32_t M, double* out) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->have_ref && M == ctx->M && out, "frx_get_reference: no reference set / wrong length");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(out, sizeof(double) * M, ctx->ref.p, sizeof(double) * ctx->Mpad, sizeof(double) * M, 6,
                         cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FRX_OK;
}

int frx_initial_state(frx_ctx* ctx, const double* x0, int32_t low_vel_mode, double wheelbase, double* x_cl) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->have_ref, "frx_initial_state: set the reference path first");
    REQUIRE(x0 && x_cl && wheelbase > 0, "frx_initial_state: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->gout.reserve(16));
    CK(cudaMemcpyAsync(ctx->gout.p, x0, sizeof(double) * 6, cudaMemcpyHostToDevice, ctx->stream));
    frx_launch_initial_state(ctx->M, ctx->Mpad, ctx->ref.p, ctx->gout.p, wheelbase, low_vel_mode ? 1 : 0, ctx->gout.p + 8, ctx->stream);
    CK(cudaGetLastError());
    double h[7];
    CK(cudaMemcpyAsync(h, ctx->gout.p + 8, sizeof(double) * 7, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 6; ++k) x_cl[k] = h[k];
    REQUIRE(h[6] == 0.0, "Initial state or reference incorrect! Curvilinear velocity is negative which indicates that the ego "
                         "vehicle is not driving in the same direction as specified by the reference");
    return FRX_OK;
}

int frx_set_params(frx_ctx* ctx, const frx_params* p) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(p != nullptr, "frx_set_params: null");
    REQUIRE(p->N >= 1 && p->N <= 63, "frx_set_params: N must be in [1, 63]");
    REQUIRE(p->dt > 0, "frx_set_params: dt must be > 0");
    REQUIRE(p->n_costs >= 0 && p->n_costs <= FRX_MAX_COSTS, "frx_set_params: n_costs out of range");
    for (int k = 0; k < p->n_costs; ++k)
        REQUIRE(p->cost_ids[k] >= 0 && p->cost_ids[k] < FRX_NUM_COST_TERMS, "frx_set_params: unknown cost id");
    REQUIRE(!p->curvature_rate_from_v_delta || p->v_delta_max > 0, "frx_set_params: v_delta_max must be > 0");
    REQUIRE(p->velocity_offset_norm >= 0 && p->velocity_offset_norm <= 2, "frx_set_params: velocity_offset_norm must be 0, 1 or 2");
    REQUIRE(p->prediction_cost_mode == 0 || p->prediction_cost_mode == 1, "frx_set_params: prediction_cost_mode must be 0 or 1");
    if (ctx->have_params && ctx->prm.N != p->N) ctx->have_tables = false;
    ctx->prm = *p;
    ctx->kappa_max = tan(p->delta_max) / p->wheelbase;   // reactive_planner.py:492
    ctx->have_params = true;
    return FRX_OK;
}

int frx_set_time_tables(frx_ctx* ctx, int32_t nT, const double* T_values, const int32_t* traj_len, const double* tpow) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->have_params, "frx_set_time_tables: call frx_set_params first");
    REQUIRE(nT >= 1 && nT <= FRX_MAX_T_VALUES && T_values && traj_len && tpow, "frx_set_time_tables: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int Nt = ctx->prm.N + 1;
    const int tpitch = nchunk_for(Nt) * 32;
    for (int k = 0; k < nT; ++k)
        REQUIRE(traj_len[k] >= 1 && traj_len[k] <= Nt, "frx_set_time_tables: traj_len exceeds the planning horizon");
    // The rounded powers of step i do not depend on the duration: np.round(np.arange(0, T + dt, dt), 5) is 0 + i * dt for
    // every T (reactive_planner.py:296-300), only the number of samples differs.  The kernels therefore keep ONE table
    // [5][tpitch] per CTA in shared memory; a caller whose tables disagree on a common step is refused.
    std::vector<double> h((size_t)5 * tpitch, 0.0);
    std::vector<int> have(Nt, 0);
    for (int k = 0; k < nT; ++k)
        for (int i = 0; i < traj_len[k]; ++i) {
            for (int p = 0; p < 5; ++p) {
                const double v = tpow[((size_t)k * 5 + p) * Nt + i];
                double& m = h[(size_t)p * tpitch + i];
                REQUIRE(!have[i] || m == v, "frx_set_time_tables: the time-power tables of two durations differ on a common step");
                m = v;
            }
            have[i] = 1;
        }
    CK(ctx->Ttab.reserve(nT)); CK(ctx->Tlen.reserve(nT)); CK(ctx->tpow.reserve(h.size()));
    CK(cudaMemcpyAsync(ctx->Ttab.p, T_values, sizeof(double) * nT, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->Tlen.p, traj_len, sizeof(int) * nT, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->tpow.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nT = nT; ctx->tpitch = tpitch; ctx->have_tables = true;
    return FRX_OK;
}

int frx_set_predictions(frx_ctx* ctx, int32_t O, int32_t T, const double* pos, const double* cov, const double* theta,
                        const double* half_len, const double* half_wid, const int32_t* len_valid) {
    if (!ctx) return FRX_ERR_INVALID;
    if (O <= 0) { ctx->O = 0; return FRX_OK; }
    REQUIRE(T >= 1 && pos && cov && theta && half_len && half_wid && len_valid, "frx_set_predictions: bad arguments");
    for (int o = 0; o < O; ++o) REQUIRE(len_valid[o] >= 0 && len_valid[o] <= T, "frx_set_predictions: len_valid out of range");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)O * T;
    CK(ctx->raw_pos.reserve(n * 2)); CK(ctx->raw_cov.reserve(n * 4)); CK(ctx->raw_theta.reserve(n));
    CK(ctx->raw_hl.reserve(O)); CK(ctx->raw_hw.reserve(O)); CK(ctx->obs_len.reserve(O));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->raw_pos.p, pos, n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->raw_cov.p, cov, n * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->raw_theta.p, theta, n * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->raw_hl.p, half_len, O * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->raw_hw.p, half_wid, O * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->obs_len.p, len_valid, O * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // host buffers may be pageable and reused by the caller
    // the SoA table (inverse covariances, hulls) is built at plan time, when the step pitch (32 x chunks of the
    // planning horizon) is known
    ctx->O = O; ctx->T = T; ctx->Tp = 0; ctx->prob_ready = false;
    return FRX_OK;
}


int frx_set_obstacle_positions(frx_ctx* ctx, int32_t n, const double* pos_xy) {
    if (!ctx) return FRX_ERR_INVALID;
    if (n <= 0) { ctx->n_obs_pos = 0; return FRX_OK; }
    REQUIRE(pos_xy != nullptr, "frx_set_obstacle_positions: null");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->obs_pos.reserve((size_t)n * 2));
    CK(cudaMemcpyAsync(ctx->obs_pos.p, pos_xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_obs_pos = n;
    return FRX_OK;
}

int frx_set_static_obbs(frx_ctx* ctx, int32_t B, const double* obb) {
    if (!ctx) return FRX_ERR_INVALID;
    if (B <= 0) { ctx->B = 0; return FRX_OK; }
    REQUIRE(obb != nullptr, "frx_set_static_obbs: null");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->raw_sobb.reserve((size_t)B * 5)); CK(ctx->sobb.reserve((size_t)B * 8));
    CK(cudaMemcpyAsync(ctx->raw_sobb.p, obb, sizeof(double) * 5 * B, cudaMemcpyHostToDevice, ctx->stream));
    frx_launch_static_prep(B, ctx->raw_sobb.p, ctx->sobb.p, ctx->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->B = B; ctx->sobb32_dirty = true;
    return FRX_OK;
}

// ---- one plan = prepare (buffers + kernel arguments) -> launch -> finish (arg-min, read-back of the result)
static int prepare_plan(frx_ctx* ctx, long long N, const double* d_sampling, bool grid_mode, int g_nv, int g_nd,
                        const double* d_t1, const double* d_v1, const double* d_d1, const double* xcl,
                        long long row_first, long long row_base, int max_grid, int seg_hint, cudaStream_t st,
                        FrxKernelArgs* a_out, int* grid_out, int* Nt_out) {
    REQUIRE(ctx->have_params && ctx->have_ref && ctx->have_tables,
            "frx_plan: frx_set_params, frx_set_reference and frx_set_time_tables must be called first");
    const frx_params& p = ctx->prm;
    const int Nt = p.N + 1, Ntp = pitch_for(Nt), nchunk = nchunk_for(Nt), K = p.n_costs;
    const long long Np = (N + 31) & ~31LL;         // candidates per (field, step) plane, padded to whole tiles
    // without store_states only the x, y, theta planes exist, and only when the obstacle pass re-reads them
    const bool need_xyt = (ctx->O > 0 || ctx->B > 0 || ctx->n_obs_pos > 0);
    if (p.store_states) CK(ctx->states.reserve((size_t)FRX_NUM_FIELDS * Nt * Np));
    else if (need_xyt) CK(ctx->states.reserve((size_t)3 * Nt * Np));
    CK(ctx->costs.reserve((size_t)N * (K > 0 ? K : 1))); CK(ctx->total.reserve(N)); CK(ctx->flags.reserve(N));
    CK(ctx->traj_len.reserve(N));
    if (ctx->occ_Mpad != ctx->Mpad || ctx->occ_Nt != Nt) {
        int b = 1;
        REQUIRE(frx_eval_smem_bytes(ctx->Mpad, Nt) <= (size_t)ctx->max_smem_optin,
                "frx_plan: reference path too long for the shared-memory table at this planning horizon");
        CK(frx_eval_occupancy(ctx->Mpad, Nt, &b));
        REQUIRE(b >= 1, "frx_plan: eval kernel does not fit on an SM with this reference length");
        ctx->occ_blocks = b; ctx->occ_Mpad = ctx->Mpad; ctx->occ_Nt = Nt;
    }
    if (ctx->O > 0 && ctx->Tp != nchunk * 32) {      // (re)build the obstacle table for this horizon
        const int Tp = nchunk * 32;
        CK(ctx->obs.reserve((size_t)ctx->O * FRX_OBS_NARR * Tp));
        CK(cudaMemsetAsync(ctx->obs.p, 0, (size_t)ctx->O * FRX_OBS_NARR * Tp * sizeof(double), st));
        frx_launch_obstacle_prep(ctx->O, ctx->T, Tp, ctx->raw_pos.p, ctx->raw_cov.p, ctx->raw_theta.p, ctx->raw_hl.p,
                                 ctx->raw_hw.p, ctx->obs.p, st);
        CK(cudaGetLastError());
        ctx->Tp = Tp;
        ctx->compact_Nt = 0;
    }
    if (ctx->O > 0 && ctx->compact_Nt != Nt) {        // per-step compact records (depend on the horizon through min(Nt, len))
        const int Tp = ctx->Tp;
        CK(ctx->opred.reserve((size_t)Tp * ctx->O * 6 + 2 * 6));   // + 2 records: frx_pred_step reads one half group ahead
        CK(ctx->ohull.reserve((size_t)Tp * ctx->O * 8));
        CK(ctx->ohull32.reserve((size_t)Tp * ctx->O));
        CK(ctx->on_pred.reserve(Tp)); CK(ctx->on_hull.reserve(Tp));
        frx_launch_obstacle_compact(ctx->O, Tp, Nt, ctx->obs.p, ctx->obs_len.p, ctx->origin_x, ctx->origin_y, ctx->opred.p,
                                    ctx->ohull.p, ctx->ohull32.p, ctx->on_pred.p, ctx->on_hull.p, st);
        CK(cudaGetLastError());
        ctx->compact_Nt = Nt;
        ctx->prob_ready = false;
    }
    if (ctx->O > 0 && p.prediction_cost_mode == 1 && !ctx->prob_ready) {
        CK(ctx->oprob.reserve((size_t)ctx->Tp * ctx->O * 8));
        frx_launch_prob_records(ctx->O, ctx->T, ctx->Tp, ctx->raw_pos.p, ctx->raw_cov.p, ctx->raw_theta.p, ctx->raw_hl.p,
                                ctx->obs_len.p, ctx->oprob.p, st);
        CK(cudaGetLastError());
        ctx->prob_ready = true;
    }
    if (ctx->B > 0 && ctx->sobb32_dirty) {
        CK(ctx->sobb32.reserve((size_t)ctx->B));
        frx_launch_static_cull(ctx->B, ctx->sobb.p, ctx->origin_x, ctx->origin_y, ctx->sobb32.p, st);
        CK(cudaGetLastError());
        ctx->sobb32_dirty = false;
    }
    const int seg = (seg_hint > 0) ? seg_hint : frx_pick_seg(N, ctx->sm_count);
    const long long n_tiles = (N + 32 / seg - 1) / (32 / seg);     // one warp per tile of 32 / seg rows
    // All resident CTA slots are used as soon as there is a tile per CTA: first tiles are dealt warp-major (tile =
    // warp-in-CTA x grid + CTA), so a plan with fewer tiles than warps leaves the idle warps spread evenly over the SMs
    // instead of filling some SMs with two busy CTAs and others with one.
    long long want = n_tiles;
    long long full = (max_grid > 0) ? max_grid : (long long)ctx->sm_count * ctx->occ_blocks;
    int grid = (int)(want < full ? want : full);
    if (grid < 1) grid = 1;
    CK(ctx->blockbest.reserve(grid));
    CK(ctx->blockcnt.reserve((size_t)grid * (CNT_REASON1 + 10)));

    FrxKernelArgs a;
    memset(&a, 0, sizeof(a));
    a.dt = p.dt; a.a_max = p.a_max; a.v_switch = p.v_switch; a.kappa_max = ctx->kappa_max; a.wb_rear = p.wb_rear_axle;
    a.inv_dt = 1.0 / p.dt; a.inv_Nt = 1.0 / (double)Nt; a.inv_step = ctx->inv_step;
    a.half_len = p.length / 2; a.half_wid = p.width / 2; a.x0_orientation = p.x0_orientation; a.v_des = p.desired_velocity;
    for (int k = 0; k < K; ++k) { a.w[k] = p.cost_weights[k]; a.cost_ids[k] = p.cost_ids[k]; }
    a.n_costs = K; a.Nt = Nt; a.Ntp = Ntp; a.low = p.low_vel_mode; a.draw = p.draw_traj_set; a.debug = p.kinematic_debug;
    a.store_states = p.store_states; a.check_collisions = p.check_collisions;
    a.kd_from_v_delta = p.curvature_rate_from_v_delta ? 1 : 0; a.vo_norm2 = (p.velocity_offset_norm == 2) ? 1 : 0;
    a.v_delta_over_wb = p.v_delta_max / p.wheelbase; a.wheelbase = p.wheelbase;
    a.ref = ctx->ref.p; a.M = ctx->M; a.Mpad = ctx->Mpad;
    a.Ttab = ctx->Ttab.p; a.Tlen = ctx->Tlen.p; a.tpow = ctx->tpow.p; a.nT = ctx->nT; a.tpitch = ctx->tpitch;
    a.mpitch = frx_memo_pitch_host(Nt);
    a.obs = ctx->obs.p; a.obs_len = ctx->obs_len.p; a.O = ctx->O; a.Tp = ctx->Tp;
    a.opred = ctx->opred.p; a.ohull = ctx->ohull.p; a.on_pred = ctx->on_pred.p; a.on_hull = ctx->on_hull.p;
    a.oprob = ctx->oprob.p; a.pred_mode = (p.prediction_cost_mode == 1) ? 1 : 0;
    a.ohull32 = ctx->ohull32.p; a.sobb32 = ctx->sobb32.p; a.origin_x = ctx->origin_x; a.origin_y = ctx->origin_y;
    a.obs_pos = ctx->obs_pos.p; a.n_obs_pos = ctx->n_obs_pos; a.sobb = ctx->sobb.p; a.B = ctx->B;
    a.sampling = grid_mode ? nullptr : d_sampling;
    a.g_t1 = d_t1; a.g_v1 = d_v1; a.g_d1 = d_d1; a.g_nv = g_nv; a.g_nd = g_nd;
    if (xcl) for (int k = 0; k < 6; ++k) a.xcl[k] = xcl[k];
    a.row_first = row_first; a.row_base = row_base; a.N = N;
    a.states = ctx->states.p; a.costs = ctx->costs.p; a.total = ctx->total.p; a.flags = ctx->flags.p;
    a.seg = seg; a.Np = Np; a.nf_store = p.store_states ? FRX_NUM_FIELDS : 3; a.keep_xyt = (!p.store_states && need_xyt) ? 1 : 0;
    a.traj_len = ctx->traj_len.p; a.blockbest = ctx->blockbest.p; a.blockcnt = ctx->blockcnt.p; a.counters = ctx->counters.p;
    a.winner = ctx->winner.p; a.host_res = ctx->d_res; a.n_cta = grid;
    a.xchg = nullptr; a.xchg_rank = ctx->xchg_rank; a.xchg_epoch = 0;
    if (ctx->counters_dirty) {
        CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * FRX_NUM_COUNTERS, st));
        ctx->counters_dirty = false;
    }
    *a_out = a; *grid_out = grid; *Nt_out = Nt;
    ctx->lastN = N; ctx->lastNp = Np; ctx->lastK = K; ctx->lastNt = Nt; ctx->lastNtp = Ntp;
    ctx->last_all_fields = p.store_states != 0;
    return FRX_OK;
}

// The eval kernel's last CTA already reduced the winners and wrote the result record to mapped host memory;
// only Planner._collision_counter (needs the winner first, then one pass over the flags) is a second kernel.
static int enqueue_finish(frx_ctx* ctx, long long N, long long row_base, int grid, cudaStream_t st) {
    (void)grid;
    const frx_params& p = ctx->prm;
    ctx->counted_last = p.check_collisions && (ctx->O > 0 || ctx->B > 0);
    if (ctx->counted_last) {
        ctx->last_launches += 1;
        long long cg = (N + 255) / 256;
        if (cg > (long long)ctx->sm_count * 4) cg = (long long)ctx->sm_count * 4;
        frx_launch_collision_counter(N, row_base, ctx->total.p, ctx->flags.p, ctx->winner.p, ctx->counters.p, (int)cg, st);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&ctx->h_res->counters[CNT_COLLISION_COUNTER], ctx->counters.p + CNT_COLLISION_COUNTER,
                           sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    }
    return FRX_OK;
}

static int fill_result(frx_ctx* ctx, long long N, frx_result* out) {
    memset(out, 0, sizeof(*out));
    out->n_rows = N;
    const HostResult& h = *ctx->h_res;
    if (h.counters[CNT_T_NOT_FOUND]) {
        ctx->err = "frx_plan: a sampling row uses a duration (column 1) that is missing from frx_set_time_tables";
        return FRX_ERR_INVALID;
    }
    out->argmin = h.winner.idx;
    out->min_cost = (h.winner.idx >= 0) ? h.winner.cost : INFINITY;
    out->n_in_list = (int64_t)h.counters[CNT_IN_LIST];
    out->n_feasible = (int64_t)h.counters[CNT_FEASIBLE];
    out->n_candidates = (int64_t)h.counters[CNT_CANDIDATES];
    out->n_collide = (int64_t)h.counters[CNT_COLLIDE];
    out->n_boundary = (int64_t)h.counters[CNT_BOUNDARY];
    out->collision_counter = ctx->counted_last ? (int64_t)h.counters[CNT_COLLISION_COUNTER] : 0;
    out->reason_counts[0] = (int64_t)h.counters[CNT_INFEASIBLE_IN_LIST];
    for (int q = 1; q <= 10; ++q) out->reason_counts[q] = (int64_t)h.counters[CNT_REASON1 + q - 1];
    return FRX_OK;
}

// Large plans with obstacle work: the obstacle pass, the arg-min and the result record run as a second kernel
// (frx_obstacle_kernel, see there for why).  FRX_SPLIT_OBS=0/1 forces the choice.
static int choose_obstacle_split(frx_ctx* ctx, FrxKernelArgs* a, int grid) {
    bool obs, xc;
    frx_features(*a, &obs, &xc);
    bool d2o = false;
    for (int k = 0; k < a->n_costs; ++k) d2o |= (a->cost_ids[k] == FRX_COST_DISTANCE_TO_OBSTACLES) && a->n_obs_pos > 0;
    bool split = (obs || d2o) && a->seg == 1;
    if (const char* e = getenv("FRX_SPLIT_OBS")) split = (obs || d2o) && e[0] == '1';
    if (obs && a->pred_mode == 1) split = true;      // the collision-probability cost only exists in frx_obstacle_kernel<1>
    a->defer_obs = split ? 1 : 0;
    if (split) {
        const size_t need = (size_t)frx_obstacle_pass_max_grid(ctx->sm_count);
        CK(ctx->blockbest.reserve(need > (size_t)grid ? need : (size_t)grid));
        if (a->N <= (1LL << 22)) {      // plans this small may be cut into step chunks (frx_obstacle_chunks)
            CK(ctx->obs_part.reserve(frx_obstacle_scratch_elems(a->N)));
            CK(ctx->obs_hit.reserve(frx_obstacle_scratch_elems(a->N)));
        }
    }
    a->blockbest = ctx->blockbest.p; a->obs_part = ctx->obs_part.p; a->obs_hit = ctx->obs_hit.p; a->obs_chunks = 1;
    return FRX_OK;
}

static int enqueue_plan(frx_ctx* ctx, long long N, const double* d_sampling, bool grid_mode, int g_nv, int g_nd,
                        const double* d_t1, const double* d_v1, const double* d_d1, const double* xcl,
                        long long row_first, long long row_base) {
    FrxRange range("frx:enqueue");
    cudaStream_t st = ctx->stream;
    FrxKernelArgs a;
    int grid = 1, Nt = 1;
    int rc = prepare_plan(ctx, N, d_sampling, grid_mode, g_nv, g_nd, d_t1, d_v1, d_d1, xcl, row_first, row_base, 0, 0, st,
                          &a, &grid, &Nt);
    if (rc != FRX_OK) return rc;
    ctx->counters_dirty = true;          // cleared again once the launch sequence has completed
#ifdef FRX_TRACE
    static DevBuf<unsigned long long> trace;
    const size_t n_tr = (size_t)grid * FRX_WARPS_PER_CTA * 16;
    CK(trace.reserve(n_tr));
    CK(cudaMemsetAsync(trace.p, 0, n_tr * 8, st));
    a.trace = trace.p;
#endif
    rc = choose_obstacle_split(ctx, &a, grid);
    if (rc != FRX_OK) return rc;
    if (ctx->xchg_dev) { a.xchg = ctx->xchg_dev; a.xchg_epoch = ++ctx->xchg_epoch; }
    ctx->last_launches = 1;
    CK(cudaEventRecord(ctx->evk0, st));
    CK(frx_launch_eval(a, Nt, grid, st));
    ctx->split_last = a.defer_obs != 0;
    if (a.defer_obs) {
        CK(cudaEventRecord(ctx->evkm, st));
        int n_obs_launches = 0;
        CK(frx_launch_obstacle_pass(a, ctx->sm_count, st, &n_obs_launches));
        ctx->last_launches += n_obs_launches;
    }
    CK(cudaEventRecord(ctx->evk1, st));
#ifdef FRX_TRACE
    {
        std::vector<unsigned long long> h(n_tr);
        CK(cudaMemcpyAsync(h.data(), trace.p, n_tr * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (const char* path = getenv("FRX_TRACE_FILE")) {
            if (FILE* f = fopen(path, "wb")) { fwrite(h.data(), 8, n_tr, f); fclose(f); }
        }
    }
#endif
    rc = enqueue_finish(ctx, N, row_base, grid, st);
    if (rc != FRX_OK) return rc;
    CK(cudaEventRecord(ctx->ev1, st));
    ctx->pending = true;
    return FRX_OK;
}

static int wait_plan(frx_ctx* ctx, frx_result* out) {
    FrxRange range("frx:wait");
    REQUIRE(out != nullptr, "frx_plan: null result");
    REQUIRE(ctx->pending, "frx_plan_wait: no plan in flight");
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->pending = false;
    ctx->counters_dirty = false;         // the last CTA re-armed them
    int rc = fill_result(ctx, ctx->lastN, out);
    if (rc != FRX_OK) return rc;
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->evk0, ctx->evk1)); out->eval_kernel_ms = ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1)); out->total_device_ms = ms;
    out->obstacle_kernel_ms = 0.f;
    if (ctx->split_last) { CK(cudaEventElapsedTime(&ms, ctx->evkm, ctx->evk1)); out->obstacle_kernel_ms = ms; }
    return FRX_OK;
}

static int run_plan(frx_ctx* ctx, long long N, const double* d_sampling, bool grid_mode, int g_nv, int g_nd,
                    const double* d_t1, const double* d_v1, const double* d_d1, const double* xcl,
                    long long row_first, long long row_base, frx_result* out) {
    REQUIRE(out != nullptr, "frx_plan: null result");
    int rc = enqueue_plan(ctx, N, d_sampling, grid_mode, g_nv, g_nd, d_t1, d_v1, d_d1, xcl, row_first, row_base);
    if (rc != FRX_OK) return rc;
    return wait_plan(ctx, out);
}

// Pinned (page-locked, mapped) host memory is read by the eval kernel in place: every warp prefetches the rows of its
// next tile over PCIe while it evaluates the current one, there is no staging copy.  FRX_ZEROCOPY=0 disables this.
static const double* zero_copy_pointer(const double* host) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("FRX_ZEROCOPY"); enabled = (e && e[0] == '0') ? 0 : 1; }
    if (!enabled) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
    return (const double*)at.devicePointer;
}

int frx_plan(frx_ctx* ctx, int64_t N, const double* sampling, int64_t row_index_base, frx_result* out) {
    if (!ctx) return FRX_ERR_INVALID;
    FrxRange range("frx_plan");
    REQUIRE(N >= 1 && sampling != nullptr, "frx_plan: empty sampling matrix");
    CK(cudaSetDevice(ctx->device));
    if (const double* mapped = zero_copy_pointer(sampling)) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        return run_plan(ctx, N, mapped, false, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, row_index_base, out);
    }
    CK(ctx->sampling.reserve((size_t)N * 13));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    CK(cudaMemcpyAsync(ctx->sampling.p, sampling, (size_t)N * 13 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return run_plan(ctx, N, ctx->sampling.p, false, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, row_index_base, out);
}

int frx_plan_device(frx_ctx* ctx, int64_t N, const void* d_sampling, int64_t row_index_base, frx_result* out) {
    if (!ctx) return FRX_ERR_INVALID;
    FrxRange range("frx_plan_device");
    REQUIRE(N >= 1 && d_sampling != nullptr, "frx_plan_device: empty sampling matrix");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return run_plan(ctx, N, (const double*)d_sampling, false, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, row_index_base, out);
}

int frx_plan_device_async(frx_ctx* ctx, int64_t N, const void* d_sampling, int64_t row_index_base) {
    if (!ctx) return FRX_ERR_INVALID;
    FrxRange range("frx_plan_device_async");
    REQUIRE(N >= 1 && d_sampling != nullptr, "frx_plan_device_async: empty sampling matrix");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return enqueue_plan(ctx, N, (const double*)d_sampling, false, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, row_index_base);
}

int frx_plan_wait(frx_ctx* ctx, frx_result* out) {
    if (!ctx) return FRX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    return wait_plan(ctx, out);
}

int frx_plan_grid(frx_ctx* ctx, int32_t nt, const double* t1, int32_t nv, const double* ss1, int32_t nd,
                  const double* d1, const double* x_cl, int64_t row_first, int64_t row_count, frx_result* out) {
    if (!ctx) return FRX_ERR_INVALID;
    FrxRange range("frx_plan_grid");
    REQUIRE(nt >= 1 && nv >= 1 && nd >= 1 && t1 && ss1 && d1 && x_cl, "frx_plan_grid: bad arguments");
    const long long total = (long long)nt * nv * nd;
    REQUIRE(row_first >= 0 && row_count >= 1 && row_first + row_count <= total, "frx_plan_grid: row range outside the grid");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->grid.reserve((size_t)nt + nv + nd));
    // the three axes travel as ONE copy out of a pinned staging buffer (three small pageable copies cost ~5 us each)
    const size_t n_axes = (size_t)nt + nv + nd;
    if (ctx->grid_stage_cap < n_axes) {
        if (ctx->grid_stage) cudaFreeHost(ctx->grid_stage);
        ctx->grid_stage = nullptr; ctx->grid_stage_cap = 0;
        CK(cudaHostAlloc((void**)&ctx->grid_stage, sizeof(double) * n_axes, cudaHostAllocDefault));
        ctx->grid_stage_cap = n_axes;
    }
    CK(cudaStreamSynchronize(ctx->stream));      // the previous plan's copy has left the staging buffer
    memcpy(ctx->grid_stage, t1, sizeof(double) * nt);
    memcpy(ctx->grid_stage + nt, ss1, sizeof(double) * nv);
    memcpy(ctx->grid_stage + nt + nv, d1, sizeof(double) * nd);
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    CK(cudaMemcpyAsync(ctx->grid.p, ctx->grid_stage, sizeof(double) * n_axes, cudaMemcpyHostToDevice, ctx->stream));
    return run_plan(ctx, row_count, nullptr, true, nv, nd, ctx->grid.p, ctx->grid.p + nt, ctx->grid.p + nt + nv, x_cl,
                    row_first, row_first, out);
}

int frx_plan_batched(int32_t n_agents, frx_ctx** ctxs, const int64_t* n_rows, const double* const* samplings,
                     frx_result* results) {
    if (n_agents < 1 || !ctxs || !n_rows || !samplings || !results || !ctxs[0]) return FRX_ERR_INVALID;
    FrxRange range("frx_plan_batched");
    frx_ctx* ctx = ctxs[0];                       // the batch runs on the first context's stream
    REQUIRE(n_agents <= 256, "frx_plan_batched: at most 256 agents per launch");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<FrxKernelArgs> args(n_agents);
    std::vector<int> cta_begin(n_agents + 1, 0), grids(n_agents, 1);
    long long total_rows = 0;
    int nchunk0 = -1, max_Mpad = 0, occ = 1;
    for (int a = 0; a < n_agents; ++a) {
        REQUIRE(ctxs[a] && ctxs[a]->device == ctx->device, "frx_plan_batched: all contexts must live on one device");
        REQUIRE(ctxs[a]->prm.prediction_cost_mode == 0, "frx_plan_batched: prediction_cost_mode 1 is not available in the batched launch");
        REQUIRE(n_rows[a] >= 1 && samplings[a], "frx_plan_batched: empty sampling matrix");
        total_rows += n_rows[a];
    }
    CK(cudaEventRecord(ctx->ev0, st));
    for (int a = 0; a < n_agents; ++a) {
        frx_ctx* c = ctxs[a];
        CK(c->sampling.reserve((size_t)n_rows[a] * 13));
        CK(cudaMemcpyAsync(c->sampling.p, samplings[a], (size_t)n_rows[a] * 13 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    // CTA budget: the persistent grid of one launch, split over the agents in proportion to their rows
    {
        int b = 1;
        for (int a = 0; a < n_agents; ++a) if (ctxs[a]->Mpad > max_Mpad) max_Mpad = ctxs[a]->Mpad;
        CK(frx_eval_occupancy(max_Mpad, ctx->prm.N + 1, &b));
        REQUIRE(b >= 1, "frx_plan_batched: eval kernel does not fit on an SM");
        occ = b;
    }
    const long long budget = (long long)ctx->sm_count * occ;
    const int batch_seg = frx_pick_seg(total_rows, ctx->sm_count);     // one kernel instance serves all agents
    for (int a = 0; a < n_agents; ++a) {
        frx_ctx* c = ctxs[a];
        long long share = (budget * n_rows[a] + total_rows - 1) / total_rows;
        if (share < 1) share = 1;
        int g = 1, nch = 1;
        int rc = prepare_plan(c, n_rows[a], c->sampling.p, false, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, 0,
                              (int)share, batch_seg, st, &args[a], &g, &nch);
        if (rc != FRX_OK) { ctx->err = c->err; return rc; }
        if (nchunk0 < 0) nchunk0 = nch;
        REQUIRE(nch == nchunk0, "frx_plan_batched: all agents must share the planning horizon (samples per candidate)");
        grids[a] = g;
        cta_begin[a + 1] = cta_begin[a] + g;
        c->counters_dirty = true;
    }
    CK(ctx->batch_args.reserve(n_agents)); CK(ctx->batch_cta.reserve(n_agents + 1));
    CK(cudaMemcpyAsync(ctx->batch_args.p, args.data(), sizeof(FrxKernelArgs) * n_agents, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->batch_cta.p, cta_begin.data(), sizeof(int) * (n_agents + 1), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->evk0, st));
    CK(frx_launch_eval_batched(args.data(), ctx->batch_args.p, ctx->batch_cta.p, n_agents, max_Mpad, nchunk0, cta_begin[n_agents], st));
    for (int a = 0; a < n_agents; ++a) ctxs[a]->split_last = false;
    // (the batch keeps the fused obstacle pass: one obstacle kernel per agent, each far below a full wave, measured
    // twice as slow -- 1.31 ms against 0.66 ms for 6 x 50,000 rows)
    CK(cudaEventRecord(ctx->evk1, st));
    for (int a = 0; a < n_agents; ++a) {
        ctxs[a]->last_launches = (a == 0) ? 1 : 0;       // the one batched eval kernel is booked on the first context
        int rc = enqueue_finish(ctxs[a], n_rows[a], 0, grids[a], st);
        if (rc != FRX_OK) { ctx->err = ctxs[a]->err; return rc; }
    }
    CK(cudaEventRecord(ctx->ev1, st));
    CK(cudaStreamSynchronize(st));     // pageable host matrices may be reused by the caller after return
    for (int a = 0; a < n_agents; ++a) ctxs[a]->counters_dirty = false;
    float kms = 0.f, tms = 0.f;
    CK(cudaEventElapsedTime(&kms, ctx->evk0, ctx->evk1));
    CK(cudaEventElapsedTime(&tms, ctx->ev0, ctx->ev1));
    for (int a = 0; a < n_agents; ++a) {
        int rc = fill_result(ctxs[a], n_rows[a], &results[a]);
        if (rc != FRX_OK) { ctx->err = ctxs[a]->err; return rc; }
        results[a].eval_kernel_ms = kms; results[a].total_device_ms = tms;     // one launch serves all agents
    }
    return FRX_OK;
}

int32_t frx_state_pitch(const frx_ctx* ctx) { return ctx ? ctx->lastNtp : 0; }
int32_t frx_last_launches(const frx_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

int frx_get_states(frx_ctx* ctx, int64_t n_idx, const int64_t* idx, uint32_t field_mask, double* out) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->lastN > 0 && ctx->last_all_fields, "frx_get_states: no materialised states (store_states = 0 or no plan yet)");
    REQUIRE(n_idx >= 1 && idx && out && field_mask && field_mask < (1u << FRX_NUM_FIELDS), "frx_get_states: bad arguments");
    for (int64_t k = 0; k < n_idx; ++k) REQUIRE(idx[k] >= 0 && idx[k] < ctx->lastN, "frx_get_states: row index out of range");
    CK(cudaSetDevice(ctx->device));
    const int nf = __builtin_popcount(field_mask);
    const size_t n_out = (size_t)nf * n_idx * ctx->lastNtp;
    CK(ctx->gidx.reserve(n_idx)); CK(ctx->gout.reserve(n_out));
    CK(cudaMemcpyAsync(ctx->gidx.p, idx, sizeof(long long) * n_idx, cudaMemcpyHostToDevice, ctx->stream));
    frx_launch_gather(ctx->states.p, ctx->lastNp, ctx->lastNt, ctx->lastNtp, ctx->gidx.p, 0, n_idx, field_mask, ctx->gout.p,
                      ctx->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->gout.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FRX_OK;
}

int frx_winner_states(frx_ctx* ctx, uint32_t field_mask, double* out) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->lastN > 0 && ctx->last_all_fields, "frx_winner_states: no materialised states (store_states = 0 or no plan yet)");
    REQUIRE(!ctx->pending, "frx_winner_states: a plan is still in flight (frx_plan_wait first)");
    REQUIRE(out && field_mask && field_mask < (1u << FRX_NUM_FIELDS), "frx_winner_states: bad arguments");
    REQUIRE(ctx->h_res->winner.idx >= 0, "frx_winner_states: the last plan selected no candidate");
    const int pitch = ctx->lastNtp, Nt = ctx->lastNt;
    int fo = 0;
    for (int f = 0; f < FRX_NUM_FIELDS; ++f) {
        if (!(field_mask & (1u << f))) continue;
        for (int i = 0; i < pitch; ++i) out[(size_t)fo * pitch + i] = (i < Nt) ? ctx->h_res->winner_states[f][i] : 0.0;
        ++fo;
    }
    return FRX_OK;
}

int frx_winner_record(frx_ctx* ctx, uint32_t* flags, int32_t* traj_len, double* total, double* costs) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->lastN > 0 && !ctx->pending, "frx_winner_record: no finished plan");
    REQUIRE(ctx->h_res->winner.idx >= 0, "frx_winner_record: the last plan selected no candidate");
    if (flags) *flags = ctx->h_res->winner_flags;
    if (traj_len) *traj_len = ctx->h_res->winner_traj_len;
    if (total) *total = ctx->h_res->winner.cost;
    if (costs) for (int k = 0; k < ctx->lastK; ++k) costs[k] = ctx->h_res->winner_costs[k];
    return FRX_OK;
}

int frx_get_states_range(frx_ctx* ctx, int64_t first, int64_t count, uint32_t field_mask, double* out) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->lastN > 0 && ctx->last_all_fields, "frx_get_states_range: no materialised states");
    REQUIRE(first >= 0 && count >= 1 && first + count <= ctx->lastN && out && field_mask && field_mask < (1u << FRX_NUM_FIELDS),
            "frx_get_states_range: bad arguments");
    CK(cudaSetDevice(ctx->device));
    // the tensor is [field][step][candidate]; the caller gets [field][candidate][pitch]: transpose on the device in
    // slabs of bounded size, then copy out
    const int nf = __builtin_popcount(field_mask);
    const size_t per_row = (size_t)ctx->lastNtp;
    long long slab = (long long)((size_t)(64u << 20) / (per_row * sizeof(double)));     // <= 64 Mi doubles... per field
    if (slab < 1) slab = 1;
    if (slab > count) slab = count;
    CK(ctx->gout.reserve((size_t)slab * per_row));
    uint32_t m = field_mask;
    for (int fo = 0; fo < nf; ++fo) {
        const int f = __builtin_ctz(m); m &= m - 1;
        for (long long done = 0; done < count; done += slab) {
            const long long n = (count - done < slab) ? (count - done) : slab;
            frx_launch_gather(ctx->states.p, ctx->lastNp, ctx->lastNt, ctx->lastNtp, nullptr, first + done, n, 1u << f, ctx->gout.p,
                              ctx->stream);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(out + ((size_t)fo * count + done) * per_row, ctx->gout.p, (size_t)n * per_row * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return FRX_OK;
}

int frx_get_costs(frx_ctx* ctx, int64_t first, int64_t count, double* costs, double* total) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->lastN > 0 && first >= 0 && count >= 1 && first + count <= ctx->lastN, "frx_get_costs: bad range");
    CK(cudaSetDevice(ctx->device));
    if (costs && ctx->lastK > 0)
        CK(cudaMemcpyAsync(costs, ctx->costs.p + (size_t)first * ctx->lastK, sizeof(double) * count * ctx->lastK,
                           cudaMemcpyDeviceToHost, ctx->stream));
    if (total) CK(cudaMemcpyAsync(total, ctx->total.p + first, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FRX_OK;
}

int frx_get_flags(frx_ctx* ctx, int64_t first, int64_t count, uint32_t* flags, int32_t* traj_len) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->lastN > 0 && first >= 0 && count >= 1 && first + count <= ctx->lastN, "frx_get_flags: bad range");
    CK(cudaSetDevice(ctx->device));
    if (flags) CK(cudaMemcpyAsync(flags, ctx->flags.p + first, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, ctx->stream));
    if (traj_len) CK(cudaMemcpyAsync(traj_len, ctx->traj_len.p + first, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FRX_OK;
}

int frx_selftest_fdiv(frx_ctx* ctx, int64_t n, const double* a, const double* b, double* q_fdiv, double* q_ieee) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(n >= 1 && a && b && q_fdiv && q_ieee, "frx_selftest_fdiv: bad arguments");
    CK(cudaSetDevice(ctx->device));
    DevBuf<double> buf;
    CK(buf.reserve((size_t)n * 4));
    CK(cudaMemcpyAsync(buf.p, a, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(buf.p + n, b, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    frx_launch_selftest_fdiv(n, buf.p, buf.p + n, buf.p + 2 * n, buf.p + 3 * n, ctx->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(q_fdiv, buf.p + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(q_ieee, buf.p + 3 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    buf.release();
    return FRX_OK;
}

int frx_selftest_divc(frx_ctx* ctx, int64_t n, const double* a, double b, double* q_divc, double* q_ieee) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(n >= 1 && a && q_divc && q_ieee, "frx_selftest_divc: bad arguments");
    CK(cudaSetDevice(ctx->device));
    DevBuf<double> buf;
    CK(buf.reserve((size_t)n * 3));
    CK(cudaMemcpyAsync(buf.p, a, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    frx_launch_selftest_divc(n, buf.p, b, buf.p + n, buf.p + 2 * n, ctx->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(q_divc, buf.p + n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(q_ieee, buf.p + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    buf.release();
    return FRX_OK;
}

static_assert(FRX_XCHG_PAGE_BYTES == FRX_EXCHANGE_PAGE_BYTES, "exchange page size");

int frx_set_exchange(frx_ctx* ctx, void* page, int32_t rank, int32_t world) {
    if (!ctx) return FRX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->xchg_registered) { cudaHostUnregister(ctx->xchg_host); ctx->xchg_registered = false; }
    ctx->xchg_host = ctx->xchg_dev = nullptr; ctx->xchg_epoch = 0;
    if (!page) return FRX_OK;
    REQUIRE(world >= 1 && world <= FRX_XCHG_MAX_RANKS && rank >= 0 && rank < world, "frx_set_exchange: bad rank / world size");
    cudaPointerAttributes at;
    const bool known = cudaPointerGetAttributes(&at, page) == cudaSuccess && at.type == cudaMemoryTypeHost;
    if (!known) {                        // (another context of this process may have registered the page already)
        cudaGetLastError();
        CK(cudaHostRegister(page, FRX_XCHG_PAGE_BYTES, cudaHostRegisterMapped | cudaHostRegisterPortable));
        ctx->xchg_registered = true;
    }
    void* d = nullptr;
    CK(cudaHostGetDevicePointer(&d, page, 0));
    ctx->xchg_host = (FrxXchgSlot*)page; ctx->xchg_dev = (FrxXchgSlot*)d;
    ctx->xchg_rank = rank; ctx->xchg_world = world;
    return FRX_OK;
}

int frx_exchange_wait(frx_ctx* ctx, int64_t timeout_us, double* min_cost, int64_t* global_row, int32_t* owner_rank) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(ctx->xchg_host && ctx->xchg_epoch > 0, "frx_exchange_wait: no exchange attached / no plan issued");
    REQUIRE(!ctx->pending, "frx_exchange_wait: wait for the plan first (frx_plan_wait)");
    const unsigned long long e = ctx->xchg_epoch;
    volatile FrxXchgSlot* slots = ctx->xchg_host + (e & 1ULL) * FRX_XCHG_MAX_RANKS;
    timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < ctx->xchg_world; ++r) {
        unsigned spins = 0;
        while (__atomic_load_n(&slots[r].epoch, __ATOMIC_ACQUIRE) != e) {
            if ((++spins & 1023u) == 0) {
                timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
                const int64_t us = (int64_t)(t1.tv_sec - t0.tv_sec) * 1000000 + (t1.tv_nsec - t0.tv_nsec) / 1000;
                if (timeout_us > 0 && us > timeout_us) {
                    ctx->err = "frx_exchange_wait: timed out waiting for rank " + std::to_string(r) + " (epoch " + std::to_string(e) + ")";
                    return FRX_ERR_CUDA;
                }
            }
        }
    }
    double bc = INFINITY; long long bi = -1; int owner = -1;
    for (int r = 0; r < ctx->xchg_world; ++r) {
        const double c = slots[r].cost; const long long i = slots[r].idx;
        if (i >= 0 && (bi < 0 || c < bc || (c == bc && i < bi))) { bc = c; bi = i; owner = r; }
    }
    if (min_cost) *min_cost = bc;
    if (global_row) *global_row = bi;
    if (owner_rank) *owner_rank = owner;
    return FRX_OK;
}

int frx_selftest_fp64_peak(frx_ctx* ctx, double* tflops) {
    if (!ctx) return FRX_ERR_INVALID;
    REQUIRE(tflops != nullptr, "frx_selftest_fp64_peak: null");
    CK(cudaSetDevice(ctx->device));
    cudaError_t e = cudaSuccess;
    *tflops = frx_measure_fp64_peak(ctx->sm_count, ctx->stream, &e);
    CK(e);
    return FRX_OK;
}

int frx_winner_device_pointer(frx_ctx* ctx, void** winner) {
    if (!ctx || !winner) return FRX_ERR_INVALID;
    *winner = ctx->winner.p;
    return FRX_OK;
}

int frx_device_pointers(frx_ctx* ctx, void** states, void** costs, void** total, void** flags) {
    if (!ctx) return FRX_ERR_INVALID;
    if (states) *states = ctx->states.p;
    if (costs) *costs = ctx->costs.p;
    if (total) *total = ctx->total.p;
    if (flags) *flags = ctx->flags.p;
    return FRX_OK;
}

}  // extern "C"
