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
constexpr int kStageFloats = kStageWFloats + kStageAFloats;   // ring depth: Cfg<RG>::STAGES (fused_ffma.cu)

enum Epilogue : int32_t {
    EPI_ACT = 0,   // v (+relu) -> dst ; optionally save relu mask bits
    EPI_HEAD = 1,  // EPI_ACT then y = v*y_std+y_mean (exp), m = y*sigma, d = m - data -> dst; optional output
    EPI_CHI2 = 2,  // r = v ; chi2_row += r*r (CHOL) or r*d (DENSE) ; optionally r -> dst
    EPI_BWD = 3,   // v*colscale (*ysave) (*mask) -> dst
    EPI_GRAD = 4,  // v = d lnL/d xhat -> prologue Jacobian -> d lnP/du -> global grad
    // training loss, normalised space (Auxilleryfunc, linna/util.py:1070-1088)
    EPI_LOSSHEAD = 5,  // v = yhat ; delta = (that - yhat | yhat - dhat | that - dhat) * mask -> dst, mask bits saved
    EPI_LOSSQ = 6      // q = v = delta @ Chat^-1 ; chi2_row += q*delta ; optionally g_yhat = -2 q mask /(cmd B) -> dst
};

enum StepFlags : int32_t {
    F_RELU = 1,
    F_SAVE_MASK = 2,    // save (v>0) bits at mask_off
    F_APPLY_MASK = 4,   // multiply by saved bits at mask_off
    F_ADD_SRC2 = 8,     // identity second operand: v += src2[c][row]   (needs K2 == 0)
    F_STORE_DST = 16,   // CHI2: keep r in dst for the backward pass
    F_MUL_YSAVE = 32,   // BWD: multiply by saved y (ypositive: d exp)
    F_SAVE_Y = 64,      // HEAD: save y into ybuf (ypositive + backward)
    F_OUT_VEC = 128,    // HEAD: write the selected vector (yhat / y / m) to the global output
    F_LOSS_GRAD = 256,  // LOSSQ: also emit d loss / d yhat (training); otherwise chi^2 only
    F_COT = 512         // HEAD: vector-Jacobian product -- dst receives cot[row][c] * d out / d yhat (cot = args.target) instead of d
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
    int32_t rm_off;           // >= 0: also store the output row-major at rm_base[rm_off + row*rm_ld + c] (training)
    int32_t rm_ld;
    int32_t pad_;
};

struct Program {
    int32_t n_steps;
    int32_t arena_features;  // per-CTA scratch, in features (x BM floats)
    int32_t mask_features;   // per-CTA mask scratch, in features (x RG bytes)
    int32_t in_buf;          // arena offset where the prologue writes xhat
    int32_t in_rm_off;       // >= 0: row-major copy of xhat for the weight-gradient kernel
    int32_t in_rm_ld;
    int32_t pad_[2];
    Step steps[kMaxSteps];
};

struct Consts {
    const float *x_mean, *x_std;
    const uint8_t *log10_flag;
    const float *y_mean, *y_std, *sigma;
    const float *prior_scale, *prior_shift;  // theta = t*scale + shift, t = u (gauss) or Phi(u) (flat)
    const int32_t *prior_kind;
    const float *data;
    const float *data_hat;  // normalised data vector (training loss)
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
    // training / loss programs
    const float *target;  // [n][n_out] targets in physical units
    const float *cmd;     // [n] chi2(target, data) per row (clamped), or nullptr => raw chi^2 out
    float *rm_base;       // row-major activation / gradient store for the weight-gradient kernel
    float loss_inv_B;     // 1 / (rows in the optimiser batch)
    int64_t rm_row0;      // row index of this launch's first row inside the row-major store
    int32_t delta_kind;   // 0: target - pred (chi2_M,nn)  1: target - data (chi2_M,d)  2: pred - data (chi2_nn,d)
    // indexed launch (fix-up of the rows the tensor-core kernel flagged): row r of the launch is walker row_index[r] of
    // `in` / `lnp` / `grad`, and the number of rows is read from device memory (no host synchronisation)
    const int32_t *row_index;   // or nullptr: row r is row r
    const int32_t *n_dev;       // or nullptr: args.n rows; else min(*n_dev, args.n)
    int32_t *zero_me;           // counter of the NEXT tensor-core launch, cleared by this one
};

// ---- weight-gradient / AdamW kernels --------------------------------------------------------
struct WgradLayer {
    int32_t gz_off, gz_ld;  // row-major d loss/d z   [B][gz_ld]  (offset into rm_base)
    int32_t x_off, x_ld;    // row-major layer input  [B][x_ld], column K preset to 1 (bias)
    int32_t N, K;           // out, in
    int32_t w_flat, b_flat; // offsets into the flat parameter vector (b_flat < 0: no bias)
    float gscale;           // alpha of the res-block second layer, else 1
    int32_t pad_[3];
};
struct WgradTile {
    int32_t layer, n0, k0, pad_;
};
struct AdamArgs {
    float *params, *m, *v;   // flat, reference state_dict order
    float *grads;            // flat gradient (written when !fuse, read by the stand-alone AdamW)
    float *blob;             // packed operands to refresh
    const int32_t *map_fwd;  // flat index -> blob float offset of the forward-layout copy
    const int32_t *map_bwd;  // flat index -> blob float offset of the backward-layout copy (-1: none)
    float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
    int32_t fuse;
};

// Gradient average over the ranks of a data-parallel step, read straight from the peers' memory over NVLink
// (adamw_peer_kernel): grads[r] + offset is rank r's flat gradient, pads[r] its signal pad.
struct PeerReduce {
    const float *const *grads;   // device array [world]
    uint32_t *const *pads;       // device array [world]
    int64_t offset;              // element offset of this step's gradient inside every rank's buffer
    int64_t avg_offset;          // >= 0: two-phase reduction, element offset of the averaged-gradient region of every buffer
    int32_t *ticket;             // device-local arrival counter of the two-phase form (zero between launches)
    int32_t slot;                // first 32-bit word of the signal pad used here (one word per source rank)
    int32_t world, rank;
    uint32_t token;              // sequence number of this reduction (monotonic, the same on every rank)
};

}  // namespace linna
