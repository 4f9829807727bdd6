// Host-side model state shared by the C-ABI translation units (cabi.cu, tc_f16.cu).
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/linna_b200.h"
#include "linna_device.cuh"

using namespace linna;

namespace linna {
struct TcContext;   // tensor-core likelihood path state (tc_f16.cu)
struct TgContext;   // tensor-core training path state (tg_gemm.cu)
}

static inline int pad4(int n) { return (n + 3) & ~3; }

struct OpHost {
    int kind, in, mid, out, act;
    float alpha;
    std::vector<float> w, b, w2, b2, ws;
    bool has_ws;
};

enum ProgKind { PROG_PREDICT = 0, PROG_LNP = 1, PROG_GRAD = 2, PROG_LOSS = 3, PROG_TRAIN = 4, PROG_VJP = 5, PROG_COUNT = 6 };

// One helper thread per model for the host-buffer entry points (cabi.cu: lnp_host_pipelined): it stages pieces of a
// pageable input next to the calling thread and hands finished results back to the caller's arrays.  It sleeps on a
// condition variable between calls; a call posts ONE job (a pointer to a structure on the caller's stack).
struct StageJob;
struct StageHelper {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    StageJob *job = nullptr;
    bool stop = false;
    void (*run)(StageJob *) = nullptr;
    StageHelper(int device, void (*fn)(StageJob *)) : run(fn)
    {
        th = std::thread([this, device] {
            cudaSetDevice(device);
            for (;;) {
                StageJob *j;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [this] { return stop || job; });
                    if (stop) return;
                    j = job, job = nullptr;
                }
                run(j);
            }
        });
    }
    void post(StageJob *j)
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = j;
        }
        cv.notify_one();
    }
    ~StageHelper()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_one();
        th.join();
    }
};

struct linna_model {
    int device = 0, num_sms = 0;
    int n_in = 0, n_out = 0, ypositive = 0;
    int64_t n_params = 0;
    std::vector<OpHost> ops;
    std::vector<float> x_mean, x_std, y_mean, y_std, sigma;
    std::vector<uint8_t> log10_flag;
    bool has_log10 = false, has_extra = false;
    std::vector<float> extra_w, extra_b;
    float extra_scale = 0.f;
    // likelihood
    bool has_like = false;
    std::vector<int32_t> prior_kind;
    std::vector<float> prior_scale, prior_shift, data, quad;
    int quad_kind = LINNA_QUAD_CHOL;
    float temperature = 1.f;
    // training (linna_train_setup)
    bool has_train = false;
    std::vector<float> data_hat, icov_hat;
    int max_batch = 0;
    float *rm = nullptr;        // row-major activation / gradient store
    size_t rm_floats = 0;
    WgradLayer *wg_layers_dev = nullptr;
    WgradTile *wg_tiles_dev = nullptr;
    int n_wg_tiles = 0;
    int32_t *map_fwd_dev = nullptr, *map_bwd_dev = nullptr;
    // tensor-core (tcgen05, bf16x3) training kernels: built with the training constants when the network shape is covered
    linna::TgContext *tg = nullptr;
    std::string tg_why;            // why not
    int train_path = 0;            // 0 auto (tensor core when available), 1 FP32 FFMA kernels, 2 tensor core only
    int last_train_kernel = 0;     // 1 FFMA, 2 tensor core
    uint32_t peer_token = 0;       // sequence number of the peer-memory gradient reductions (linna_train_adamw_peer)
    int32_t *peer_ticket = nullptr;   // arrival counter of the two-phase reduction
    // device state
    float *blob = nullptr;
    size_t blob_floats = 0;
    Program *prog_dev = nullptr;  // [PROG_COUNT]
    Program prog_host[PROG_COUNT];
    bool prog_valid[PROG_COUNT] = {false, false, false, false, false, false};
    Consts consts;
    float *arena = nullptr;
    uint8_t *masks = nullptr;
    size_t arena_bytes = 0, masks_bytes = 0;
    int occ[3] = {0, 0, 0};  // CTAs/SM for RG = 1, 2, 4
    int force_rows = 0;
    bool fold_enabled = true;     // fold last layer + inverse transform + Cholesky product for lnP
    // the scratch arena is shared by every launch on this model: launches on different streams are
    // chained through this event so that they never overlap
    cudaEvent_t last_done = nullptr;
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    // tensor-core (tcgen05) path, built lazily for large batches
    linna::TcContext *tc = nullptr;
    bool tc_failed = false;
    std::string tc_why;           // why the tensor-core program could not be built for this model
    int path = 0;                 // 0 auto (cluster kernel below tc_min_rows rows, tensor core from there on), 1 FFMA only,
                                  // 2 tensor core only, 3 cluster kernel only
    int last_kernel = 0;          // kernel that served the last launch: 1 FFMA, 2 tensor core, 3 cluster (small batch)
    // small-batch cluster kernel (cluster_ffma.cu): launch geometry per program, resolved at first use
    int cl_state[PROG_COUNT] = {0, 0, 0, 0, 0, 0};   // 0 not resolved yet, 1 available, -1 unavailable (cl_why)
    int cl_cs[PROG_COUNT] = {0}, cl_depth[PROG_COUNT] = {0}, cl_clusters[PROG_COUNT] = {0}, cl_chi_q[PROG_COUNT] = {0};
    size_t cl_smem[PROG_COUNT] = {0};
    std::string cl_why;
    int64_t cl_max_rows = 255;    // auto mode: batches up to this many rows (and below tc_min_rows) take the cluster kernel
    int64_t tc_min_rows = 256;    // one full walker pair; one tensor-core pass (0.14 ms at C3) beats the FFMA kernel (0.20 ms) at every size (scratch/crossover.py)
    // host-buffer API staging
    cudaStream_t hstream = nullptr;                  // compute stream of the host-buffer entry points
    cudaStream_t cstream = nullptr, dstream = nullptr;   // host->device / device->host copy streams (pipelined chunks)
    std::vector<cudaEvent_t> pipe_events;            // 3 per chunk: input landed, kernel finished, results on the host
    float *h_stage = nullptr;                        // pinned staging for results that go to pageable user buffers
    size_t h_stage_cap = 0;
    float *h_in_stage = nullptr;                     // pinned staging for pageable input buffers
    size_t h_in_stage_cap = 0;
    StageHelper *helper = nullptr;                   // second host thread of the pipelined host-buffer calls
    float *d_in = nullptr, *d_out = nullptr, *d_lnp = nullptr, *d_grad = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0, d_lnp_cap = 0, d_grad_cap = 0;
};


// Folded tail of the likelihood: the last linear layer, the inverse output transform, the residual and
// the Cholesky product are all affine in s (the last hidden activation):
//   r = L^T (sigma*(y_std*(W s + b) + y_mean) - data) = Af s + cf ,  chi^2 = |r|^2
// Af [n_out][K] and cf [n_out] are formed in float64 once (reference arithmetic: linna/nn.py:129,
// linna/util.py:532-542, :457-458, :953-955).
static inline bool linna_can_fold(const linna_model *m)
{
    if (m->ops.size() < 2) return false;
    const OpHost &lastop = m->ops.back();
    return m->fold_enabled && m->has_like && m->quad_kind == LINNA_QUAD_CHOL && !m->ypositive && !m->has_extra &&
           lastop.kind == LINNA_OP_LINEAR && lastop.act == LINNA_ACT_NONE;
}

static inline void linna_fold_tail(const linna_model *m, std::vector<float> &Af, std::vector<float> &cff)
{
    const OpHost &lastop = m->ops.back();
    const int K = lastop.in, n_out = m->n_out;
    std::vector<double> T((size_t)n_out * K), A((size_t)n_out * K, 0.0), dvec(n_out), cf(n_out, 0.0);
    for (int j = 0; j < n_out; ++j) {
        const double sc = (double)m->sigma[j] * (double)m->y_std[j];
        for (int k = 0; k < K; ++k) T[(size_t)j * K + k] = sc * (double)lastop.w[(size_t)j * K + k];
        dvec[j] = (double)m->sigma[j] * ((double)m->y_std[j] * (double)lastop.b[j] + (double)m->y_mean[j]) -
                  (double)m->data[j];
    }
    for (int j = 0; j < n_out; ++j)            // A[n][:] += L[j][n] * T[j][:]  (L lower triangular)
        for (int n = 0; n <= j; ++n) {
            const double l = (double)m->quad[(size_t)j * n_out + n];
            if (l == 0.0) continue;
            double *a = &A[(size_t)n * K];
            const double *t = &T[(size_t)j * K];
            for (int k = 0; k < K; ++k) a[k] += l * t[k];
            cf[n] += l * dvec[j];
        }
    Af.resize((size_t)n_out * K), cff.resize(n_out);
    for (size_t i = 0; i < Af.size(); ++i) Af[i] = (float)A[i];
    for (int n = 0; n < n_out; ++n) cff[n] = (float)cf[n];
}
