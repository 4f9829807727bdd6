// Host-side model state shared by the C-ABI translation units (cabi.cu, tc_mlp.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/linna_b200.h"
#include "linna_device.cuh"

using namespace linna;

namespace linna {
struct TcContext;   // tensor-core path state (tc_mlp.cu)
}

static inline int pad4(int n) { return (n + 3) & ~3; }

struct OpHost {
    int kind, in, mid, out, act;
    float alpha;
    std::vector<float> w, b, w2, b2, ws;
    bool has_ws;
};

enum ProgKind { PROG_PREDICT = 0, PROG_LNP = 1, PROG_GRAD = 2, PROG_LOSS = 3, PROG_TRAIN = 4, PROG_COUNT = 5 };

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
    // device state
    float *blob = nullptr;
    size_t blob_floats = 0;
    Program *prog_dev = nullptr;  // [PROG_COUNT]
    Program prog_host[PROG_COUNT];
    bool prog_valid[PROG_COUNT] = {false, false, false, false, false};
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
    int path = 1;                 // 0 auto, 1 FFMA only, 2 tensor core only
    int64_t tc_min_rows = 8192;
    // host-buffer API staging
    cudaStream_t hstream = nullptr;
    float *d_in = nullptr, *d_out = nullptr, *d_lnp = nullptr, *d_grad = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0, d_lnp_cap = 0, d_grad_cap = 0;
};

