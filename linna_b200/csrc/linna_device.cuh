// Device-side data structures shared by the kernels and the C-ABI packer.
//
// The emulator likelihood is executed as a short "step program": every step is one fused
// GEMM  acc = scale*(A1 @ Wt1) + A2 @ Wt2  over a tile of BM walkers, followed by an
// epilogue (bias/relu, inverse output transform, chi^2 reduction, relu-mask backward, ...).
// Activations live feature-major ([feature][row-in-tile]) in a per-CTA scratch arena that
// stays L2 resident; weights are packed k-major ([k][n], rows padded to 4 floats) so that a
// k-slab is one coalesced cp.async stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace linna {

constexpr int kThreads = 256;        // threads per CTA of the fused FFMA kernel
constexpr int kMaxSteps = 64;
constexpr int kStageWFloats = 4096;  // weight slab per pipeline stage (16 KB)
constexpr int kStageAFloats = 2048;  // activation slab per pipeline stage (<= 8 KB)
constexpr int kStageFloats = kStageWFloats + kStageAFloats;
constexpr int kStages = 3;

enum Epilogue : int32_t {
    EPI_ACT = 0,   // v (+relu) -> dst ; optionally save relu mask bits
    EPI_HEAD = 1,  // EPI_ACT then y = v*y_std+y_mean (exp), m = y*sigma, d = m - data -> dst; optional output
    EPI_CHI2 = 2,  // r = v ; chi2_row += r*r (CHOL) or r*d (DENSE) ; optionally r -> dst
    EPI_BWD = 3,   // v*colscale (*ysave) (*mask) -> dst
    EPI_GRAD = 4   // v = d lnL/d xhat -> prologue Jacobian -> d lnP/du -> global grad
};

enum StepFlags : int32_t {
    F_RELU = 1,
    F_SAVE_MASK = 2,    // save (v>0) bits at mask_off
    F_APPLY_MASK = 4,   // multiply by saved bits at mask_off
    F_ADD_SRC2 = 8,     // identity second operand: v += src2[c][row]   (needs K2 == 0)
    F_STORE_DST = 16,   // CHI2: keep r in dst for the backward pass
    F_MUL_YSAVE = 32,   // BWD: multiply by saved y (ypositive: d exp)
    F_SAVE_Y = 64,      // HEAD: save y into ybuf (ypositive + backward)
    F_OUT_VEC = 128     // HEAD: write the selected vector (yhat / y / m) to the global output
};

struct Step {
    const float *wt1;       // [K1][ldw1]
    const float *wt2;       // [K2][ldw2] or nullptr
    const float *bias;      // [N] or nullptr   (added as scale*bias)
    const float *colscale;  // [N] or nullptr   (EPI_BWD)
    int32_t K1, K2, N, ldw1, ldw2;
    int32_t src1, src2, dst;  // arena offsets in features, -1 = none
    int32_t epi, flags;
    int32_t mask_off;         // mask arena offset in features
    int32_t ybuf;             // arena offset of the saved y (F_SAVE_Y / F_MUL_YSAVE)
    float scale;
    int32_t pad_;
};

struct Program {
    int32_t n_steps;
    int32_t arena_features;  // per-CTA scratch, in features (x BM floats)
    int32_t mask_features;   // per-CTA mask scratch, in features (x RG bytes)
    int32_t in_buf;          // arena offset where the prologue writes xhat
    Step steps[kMaxSteps];
};

struct Consts {
    const float *x_mean, *x_std;
    const uint8_t *log10_flag;
    const float *y_mean, *y_std, *sigma;
    const float *prior_scale, *prior_shift;  // theta = t*scale + shift, t = u (gauss) or Phi(u) (flat)
    const int32_t *prior_kind;
    const float *data;
    int32_t n_in, n_out, ypositive, quad_kind;
    float inv_T;
};

struct KernelArgs {
    const Program *prog;
    Consts c;
    const float *in;   // [n][n_in]   latent u (or physical theta when input_theta)
    float *out_vec;    // [n][n_out]  (predict) or nullptr
    float *lnp;        // [n] or nullptr
    float *grad;       // [n][n_in] or nullptr
    float *arena;
    uint8_t *masks;
    int64_t n;
    int32_t input_theta;  // 1: `in` holds physical parameters (Predictor.predict)
    int32_t out_kind;     // LINNA_OUT_*
};

}  // namespace linna
