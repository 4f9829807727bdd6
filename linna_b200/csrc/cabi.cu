// C-ABI of linna_b200 (see include/linna_b200.h): model packing, step-program construction and
// kernel dispatch.  Host code only; the kernels live in fused_ffma.cu, tc_f16.cu, train_kernels.cu and sampler_kernels.cu.
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "linna_host.hpp"

namespace linna {
cudaError_t launch_fused_ffma(const KernelArgs &args, int rg, int grid, cudaStream_t stream);
int fused_ffma_max_ctas_per_sm(int rg);
size_t cluster_ffma_smem_bytes(const Program &pg, int n_out, int cs, int depth, int *chi_q_out);
int cluster_ffma_max_clusters(int cs, int depth, size_t smem);
cudaError_t launch_cluster_ffma(const KernelArgs &args, int chi_q, int cs, int depth, int clusters, size_t smem, cudaStream_t stream);
int cluster_ffma_debug_read(long long *out, int max_values);
cudaError_t launch_wgrad(const WgradLayer *layers, const WgradTile *tiles, int n_tiles, const float *rm_base, int B,
                         const AdamArgs &ad, cudaStream_t stream);
cudaError_t launch_adamw(const AdamArgs &ad, int n_params, int num_sms, cudaStream_t stream);
cudaError_t launch_adamw_peer(const AdamArgs &ad, int n_params, int num_sms, const PeerReduce &pr, cudaStream_t stream);
cudaError_t launch_mean(const float *x, int n, float *out, cudaStream_t stream);
cudaError_t launch_fill_col(float *base, int ld, int col, int rows, float value, cudaStream_t stream);
cudaError_t launch_scatter_params(const float *params, float *blob, const int32_t *map_fwd, const int32_t *map_bwd,
                                  int n_params, cudaStream_t stream);
TcContext *tc_build(const linna_model *m, std::string &why);
void tc_destroy(TcContext *t);
cudaError_t tc_launch_lnp(const linna_model *m, TcContext *t, const float *u, int64_t n, float *lnp, cudaStream_t stream);
cudaError_t tc_launch_grad(const linna_model *m, TcContext *t, const float *u, int64_t n, float *lnp, float *grad,
                           cudaStream_t stream);
bool tc_has_grad(const TcContext *t);
bool tc_has_lnp(const TcContext *t);
bool tc_has_predict(const TcContext *t);
cudaError_t tc_launch_predict(const linna_model *m, TcContext *t, const float *theta, int64_t n, float *out, int out_kind,
                              cudaStream_t stream);
void tc_fix_buffers(TcContext *t, const int32_t **rows, const int32_t **count, int32_t **next_count);
void tc_launch_done(TcContext *t);
int tc_debug_read(TcContext *t, long long *out, int max_ctas);
TgContext *tg_build(const linna_model *m, const std::vector<std::array<int, 5>> &flat_off,
                    const std::vector<std::array<const float *, 2>> &bias_ptr, const std::vector<std::array<int64_t, 8>> &blob_off,
                    std::string &why);
void tg_destroy(TgContext *t);
cudaError_t tg_repack(TgContext *t, const float *params, cudaStream_t stream);
int tg_train_step(const linna_model *m, TgContext *t, const float *X, const float *Y, const float *cmd, int64_t B, const AdamArgs &ad,
                  float *loss_rows, float *loss_mean, cudaStream_t stream, bool leave_unjoined);
int tg_chisq(const linna_model *m, TgContext *t, const float *X, const float *Y, int64_t n, int kind, float *chi2, cudaStream_t stream);
int tg_debug_read(TgContext *t, long long *out, int max_steps);
}  // namespace linna

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
// The entry points work on the model's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;   // nothing to restore
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define LINNA_ON_DEVICE(m)                                                                         \
    DeviceGuard dev_guard_((m)->device);                                                           \
    if (!dev_guard_.ok) return fail(LINNA_ECUDA, "cudaSetDevice(%d) failed", (m)->device)

#define CUDA_TRY(x)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) return fail(LINNA_ECUDA, "%s: %s (%s:%d)", #x, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                       \
    } while (0)

// ----------------------------------------------------------------------------------------------
// blob builder: one host mirror, sub-allocations aligned to 64 floats (256 B)
struct Builder {
    std::vector<float> h;
    size_t alloc(size_t n)
    {
        size_t o = (h.size() + 63) / 64 * 64;
        h.resize(o + n, 0.f);
        return o;
    }
    size_t put(const std::vector<float> &v)
    {
        size_t o = alloc(v.size());
        std::copy(v.begin(), v.end(), h.begin() + o);
        return o;
    }
    // W is [out][in] (torch layout).  forward operand: Wt[k=in][n=out], ld = pad4(out)
    size_t put_fwd(const std::vector<float> &W, int out, int in, float scale = 1.f)
    {
        int ld = pad4(out);
        size_t o = alloc((size_t)in * ld);
        for (int n = 0; n < out; ++n)
            for (int k = 0; k < in; ++k) h[o + (size_t)k * ld + n] = scale * W[(size_t)n * in + k];
        return o;
    }
    // backward operand: Wb[k=out][n=in] = W itself with rows padded to pad4(in)
    size_t put_bwd(const std::vector<float> &W, int out, int in, float scale = 1.f)
    {
        int ld = pad4(in);
        size_t o = alloc((size_t)out * ld);
        for (int n = 0; n < out; ++n)
            for (int k = 0; k < in; ++k) h[o + (size_t)n * ld + k] = scale * W[(size_t)n * in + k];
        return o;
    }
};

struct OpOffsets {
    size_t w_f, w_b, b, w2_f, w2_b, b2, ws_f, ws_b;
    int mask_y, mask_h;  // mask arena offsets (features), -1 if none
};

static void free_device(linna_model *m)
{
    DeviceGuard guard(m->device);
    if (m->blob) cudaFree(m->blob);
    if (m->prog_dev) cudaFree(m->prog_dev);
    if (m->arena) cudaFree(m->arena);
    if (m->masks) cudaFree(m->masks);
    if (m->peer_ticket) { cudaFree(m->peer_ticket); m->peer_ticket = nullptr; }
    if (m->rm) cudaFree(m->rm);
    if (m->wg_layers_dev) cudaFree(m->wg_layers_dev);
    if (m->wg_tiles_dev) cudaFree(m->wg_tiles_dev);
    if (m->map_fwd_dev) cudaFree(m->map_fwd_dev);
    if (m->map_bwd_dev) cudaFree(m->map_bwd_dev);
    if (m->tg) { tg_destroy(m->tg); m->tg = nullptr; }
    m->blob = nullptr, m->prog_dev = nullptr, m->arena = nullptr, m->masks = nullptr, m->rm = nullptr;
    m->wg_layers_dev = nullptr, m->wg_tiles_dev = nullptr, m->map_fwd_dev = nullptr, m->map_bwd_dev = nullptr;
}

// (Re)build the device blob and the three step programs from the host copies.
static int rebuild(linna_model *m)
{
    LINNA_ON_DEVICE(m);
    CUDA_TRY(cudaDeviceSynchronize());  // no launch may still be reading the blob we are about to replace
    if (m->tc) { tc_destroy(m->tc); m->tc = nullptr; }
    if (m->tg) { tg_destroy(m->tg); m->tg = nullptr; }
    m->tc_failed = false;
    for (int i = 0; i < PROG_COUNT; ++i) m->cl_state[i] = 0;
    const int n_in = m->n_in, n_out = m->n_out;
    Builder B;
    std::vector<OpOffsets> off(m->ops.size());
    int maxW = std::max(n_in, n_out), maxMid = 4, mask_total = 0;
    for (size_t i = 0; i < m->ops.size(); ++i) {
        OpHost &op = m->ops[i];
        OpOffsets &o = off[i];
        maxW = std::max(maxW, std::max(op.in, op.out));
        o.mask_y = o.mask_h = -1;
        if (op.kind == LINNA_OP_LINEAR) {
            o.w_f = B.put_fwd(op.w, op.out, op.in);
            o.w_b = B.put_bwd(op.w, op.out, op.in);
            o.b = B.put(op.b);
            if (op.act == LINNA_ACT_RELU) { o.mask_y = mask_total; mask_total += op.out; }
        } else {
            maxMid = std::max(maxMid, op.mid);
            o.w_f = B.put_fwd(op.w, op.mid, op.in);
            o.w_b = B.put_bwd(op.w, op.mid, op.in);
            o.b = B.put(op.b);
            o.w2_f = B.put_fwd(op.w2, op.out, op.mid);
            o.w2_b = B.put_bwd(op.w2, op.out, op.mid);
            o.b2 = B.put(op.b2);
            if (op.has_ws) {
                o.ws_f = B.put_fwd(op.ws, op.out, op.in);
                o.ws_b = B.put_bwd(op.ws, op.out, op.in);
            }
            o.mask_h = mask_total; mask_total += op.mid;
            o.mask_y = mask_total; mask_total += op.out;
        }
    }
    size_t o_xmean = B.put(m->x_mean), o_xstd = B.put(m->x_std), o_ymean = B.put(m->y_mean), o_ystd = B.put(m->y_std);
    size_t o_sigma = B.put(m->sigma);
    size_t o_log10 = 0;
    if (m->has_log10) {
        o_log10 = B.alloc((n_in + 3) / 4);
        memcpy(&B.h[o_log10], m->log10_flag.data(), n_in);
    }
    size_t o_extra_f = 0, o_extra_b = 0, o_lastbias = 0;
    if (m->has_extra) {
        // yhat += s*(Wl xhat + bl): fold s into the packed operand and s*bl into the last bias
        o_extra_f = B.put_fwd(m->extra_w, n_out, n_in, m->extra_scale);
        o_extra_b = B.put_bwd(m->extra_w, n_out, n_in, m->extra_scale);
        OpHost &last = m->ops.back();
        std::vector<float> bb = last.b;
        for (int j = 0; j < n_out; ++j) bb[j] += m->extra_scale * m->extra_b[j];
        o_lastbias = B.put(bb);
    }
    size_t o_pk = 0, o_ps = 0, o_psh = 0, o_data = 0, o_quadF = 0, o_quadB = 0, o_cs = 0;
    if (m->has_like) {
        o_pk = B.alloc(n_in);
        memcpy(&B.h[o_pk], m->prior_kind.data(), sizeof(int32_t) * n_in);
        o_ps = B.put(m->prior_scale);
        o_psh = B.put(m->prior_shift);
        o_data = B.put(m->data);
        // chi^2 operand in [k][n] form: r_n = sum_k Q[k][n] d_k.  CHOL: Q = L (r = L^T d); DENSE: Q = C^-1.
        o_quadF = B.put_bwd(m->quad, n_out, n_out);   // Q itself, rows padded
        // backward operand: g_k = sum_n Qb[n][k] r_n with Qb = L^T stored row-major => put_fwd(L)
        o_quadB = (m->quad_kind == LINNA_QUAD_CHOL) ? B.put_fwd(m->quad, n_out, n_out) : o_quadF;
        std::vector<float> cs(n_out);
        for (int j = 0; j < n_out; ++j) cs[j] = -(m->sigma[j] * m->y_std[j]) / m->temperature;
        o_cs = B.put(cs);
    }
    // Folded tail for lnP: the last linear layer, the inverse output transform, the residual and the
    // Cholesky product are all affine in s (the last hidden activation):
    //   r = L^T (sigma*(y_std*(W s + b) + y_mean) - data) = Af s + cf ,  chi^2 = |r|^2
    // Af, cf are formed in float64 on the host once; the LNP/GRAD programs then run ONE n_out x K GEMM
    // instead of two (Predictor.predict keeps the unfolded layer: it has to return m itself).
    size_t o_foldF = 0, o_foldB = 0, o_foldc = 0, o_foldcs = 0;
    bool fold = false;
    {
        const OpHost &lastop = m->ops.back();
        fold = m->fold_enabled && m->has_like && m->quad_kind == LINNA_QUAD_CHOL && !m->ypositive && !m->has_extra &&
               lastop.kind == LINNA_OP_LINEAR && lastop.act == LINNA_ACT_NONE && m->ops.size() >= 2;
        if (fold) {
            const int K = lastop.in;
            std::vector<float> Af, cff, csf(K, -1.0f / m->temperature);
            linna_fold_tail(m, Af, cff);
            o_foldF = B.put_fwd(Af, n_out, K);
            o_foldB = B.put_bwd(Af, n_out, K);
            o_foldc = B.put(cff);
            o_foldcs = B.put(csf);
        }
    }

    size_t o_dhat = 0, o_icov = 0;
    if (m->has_train) {
        o_dhat = B.put(m->data_hat);
        o_icov = B.put_bwd(m->icov_hat, n_out, n_out);
    }

    // ---- flat parameter vector (reference state_dict order) and its maps into the packed copies
    struct FlatOff { int w = -1, b = -1, w2 = -1, b2 = -1, ws = -1; };
    std::vector<FlatOff> fo(m->ops.size());
    int64_t nflat = 0;
    for (size_t i = 0; i < m->ops.size(); ++i) {
        const OpHost &op = m->ops[i];
        if (op.kind == LINNA_OP_LINEAR) {
            fo[i].w = (int)nflat, nflat += (int64_t)op.out * op.in;
            fo[i].b = (int)nflat, nflat += op.out;
        } else {
            fo[i].w = (int)nflat, nflat += (int64_t)op.mid * op.in;
            fo[i].b = (int)nflat, nflat += op.mid;
            fo[i].w2 = (int)nflat, nflat += (int64_t)op.out * op.mid;
            fo[i].b2 = (int)nflat, nflat += op.out;
            if (op.has_ws) fo[i].ws = (int)nflat, nflat += (int64_t)op.out * op.in;
        }
    }
    std::vector<int32_t> map_fwd, map_bwd;
    if (m->has_train) {
        map_fwd.assign(nflat, -1), map_bwd.assign(nflat, -1);
        auto map_w = [&](int f0, size_t of, size_t ob, int N, int K) {
            for (int n = 0; n < N; ++n)
                for (int k = 0; k < K; ++k) {
                    map_fwd[f0 + (size_t)n * K + k] = (int32_t)(of + (size_t)k * pad4(N) + n);
                    map_bwd[f0 + (size_t)n * K + k] = (int32_t)(ob + (size_t)n * pad4(K) + k);
                }
        };
        auto map_b = [&](int f0, size_t ob, int N) { for (int n = 0; n < N; ++n) map_fwd[f0 + n] = (int32_t)(ob + n); };
        for (size_t i = 0; i < m->ops.size(); ++i) {
            const OpHost &op = m->ops[i];
            const OpOffsets &o = off[i];
            if (op.kind == LINNA_OP_LINEAR) {
                map_w(fo[i].w, o.w_f, o.w_b, op.out, op.in), map_b(fo[i].b, o.b, op.out);
            } else {
                map_w(fo[i].w, o.w_f, o.w_b, op.mid, op.in), map_b(fo[i].b, o.b, op.mid);
                map_w(fo[i].w2, o.w2_f, o.w2_b, op.out, op.mid), map_b(fo[i].b2, o.b2, op.out);
                if (op.has_ws) map_w(fo[i].ws, o.ws_f, o.ws_b, op.out, op.in);
            }
        }
    }
    // ---- row-major store layout for the weight-gradient kernel
    struct RM { int off = -1, ld = 0; };
    std::vector<RM> rm_act(m->ops.size()), rm_hid(m->ops.size()), rm_gz(m->ops.size()), rm_gzh(m->ops.size());
    size_t rm_total = 0;
    if (m->has_train) {
        auto rm_alloc = [&](int width) {
            RM r;
            r.ld = pad4(width);
            r.off = (int)rm_total;
            rm_total += ((size_t)m->max_batch * r.ld + 63) / 64 * 64;
            return r;
        };
        for (size_t i = 0; i < m->ops.size(); ++i) {
            const OpHost &op = m->ops[i];
            rm_act[i] = rm_alloc(op.in + 1);
            rm_gz[i] = rm_alloc(op.out);
            if (op.kind == LINNA_OP_RES) rm_hid[i] = rm_alloc(op.mid + 1), rm_gzh[i] = rm_alloc(op.mid);
        }
    }
    int mask_loss = mask_total;
    if (m->has_train) mask_total += n_out;

    // ---- upload
    if (m->blob && m->blob_floats < B.h.size()) { cudaFree(m->blob); m->blob = nullptr; }
    if (!m->blob) {
        m->blob_floats = B.h.size() + 1024;
        CUDA_TRY(cudaMalloc(&m->blob, m->blob_floats * sizeof(float)));
    }
    CUDA_TRY(cudaMemcpy(m->blob, B.h.data(), B.h.size() * sizeof(float), cudaMemcpyHostToDevice));
    auto P = [&](size_t o) { return (const float *)(m->blob + o); };

    Consts &c = m->consts;
    memset(&c, 0, sizeof c);
    c.x_mean = P(o_xmean), c.x_std = P(o_xstd), c.y_mean = P(o_ymean), c.y_std = P(o_ystd), c.sigma = P(o_sigma);
    c.log10_flag = m->has_log10 ? (const uint8_t *)P(o_log10) : nullptr;
    c.n_in = n_in, c.n_out = n_out, c.ypositive = m->ypositive, c.quad_kind = m->quad_kind;
    c.inv_T = 1.0f / m->temperature;
    if (m->has_like) {
        c.prior_kind = (const int32_t *)P(o_pk);
        c.prior_scale = P(o_ps), c.prior_shift = P(o_psh), c.data = P(o_data);
    }
    if (m->has_train) c.data_hat = P(o_dhat);

    // ---- arena layout (features): X | A | B | H | Y | GX
    const int bufX = 0, bufA = n_in, bufB = bufA + maxW, bufH = bufB + maxW, bufY = bufH + maxMid,
              bufGX = bufY + n_out, arena_features = bufGX + n_in;

    for (int pk = 0; pk < PROG_COUNT; ++pk) {
        m->prog_valid[pk] = false;
        if ((pk == PROG_LNP || pk == PROG_GRAD) && !m->has_like) continue;
        if ((pk == PROG_LOSS || pk == PROG_TRAIN) && !m->has_train) continue;
        const bool vjp = pk == PROG_VJP;
        if (vjp && m->has_extra) continue;            // (ChtoModelv2_linear: no vector-Jacobian program)
        // `train`: the forward and backward steps also leave row-major copies for the weight-gradient kernel
        const bool train = pk == PROG_TRAIN || (vjp && m->has_train), lossprog = pk == PROG_LOSS || pk == PROG_TRAIN;
        const bool grad = pk == PROG_GRAD || train || vjp;   // forward saves relu masks
        Program &pg = m->prog_host[pk];
        memset(&pg, 0, sizeof pg);
        pg.arena_features = arena_features;
        pg.mask_features = grad ? mask_total : 0;
        pg.in_buf = bufX;
        pg.in_rm_off = train ? rm_act[0].off : -1;
        pg.in_rm_ld = train ? rm_act[0].ld : 0;
        int ns = 0;
        auto new_step = [&]() -> Step & {
            Step &s = pg.steps[ns++];
            memset(&s, 0, sizeof s);
            s.src1 = s.src2 = s.dst = -1;
            s.scale = 1.f;
            s.mask_off = 0;
            s.rm_off = -1;
            return s;
        };
        if ((int)m->ops.size() * 2 + 2 * (int)m->ops.size() + 6 > kMaxSteps)
            return fail(LINNA_EINVAL, "too many layers (%zu)", m->ops.size());
        int cur = bufX;
        auto other = [&](int b) { return b == bufA ? bufB : bufA; };
        // ------------------------------ forward
        const bool folded = fold && (pk == PROG_LNP || pk == PROG_GRAD);
        for (size_t i = 0; i < m->ops.size(); ++i) {
            const OpHost &op = m->ops[i];
            const OpOffsets &o = off[i];
            const bool last = i + 1 == m->ops.size();
            if (last && folded) break;   // absorbed into the chi^2 step below
            if (op.kind == LINNA_OP_LINEAR) {
                Step &s = new_step();
                s.src1 = cur, s.K1 = op.in, s.wt1 = P(o.w_f), s.ldw1 = pad4(op.out), s.N = op.out;
                s.bias = P(o.b);
                if (op.act == LINNA_ACT_RELU) {
                    s.flags |= F_RELU;
                    if (grad) s.flags |= F_SAVE_MASK, s.mask_off = o.mask_y;
                }
                s.epi = last ? (lossprog ? EPI_LOSSHEAD : EPI_HEAD) : EPI_ACT;
                s.dst = other(cur);
                if (train && !last) s.rm_off = rm_act[i + 1].off, s.rm_ld = rm_act[i + 1].ld;
                if (last && lossprog && train) s.flags |= F_SAVE_MASK, s.mask_off = mask_loss;
                if (last && vjp) {
                    s.flags |= F_COT | F_OUT_VEC;
                    if (train) s.rm_off = rm_gz[i].off, s.rm_ld = rm_gz[i].ld;
                }
                if (last && m->has_extra) {
                    s.src2 = bufX, s.K2 = n_in, s.wt2 = P(o_extra_f), s.ldw2 = pad4(n_out), s.bias = P(o_lastbias);
                }
                cur = s.dst;
            } else {
                if (last && m->has_extra) return fail(LINNA_EINVAL, "extra linear branch needs a LINEAR last op");
                Step &h = new_step();
                h.src1 = cur, h.K1 = op.in, h.wt1 = P(o.w_f), h.ldw1 = pad4(op.mid), h.N = op.mid, h.bias = P(o.b);
                h.flags = F_RELU | (grad ? F_SAVE_MASK : 0), h.mask_off = o.mask_h, h.epi = EPI_ACT, h.dst = bufH;
                if (train) h.rm_off = rm_hid[i].off, h.rm_ld = rm_hid[i].ld;
                Step &y = new_step();
                y.src1 = bufH, y.K1 = op.mid, y.wt1 = P(o.w2_f), y.ldw1 = pad4(op.out), y.N = op.out;
                y.bias = P(o.b2), y.scale = op.alpha;
                y.src2 = cur;
                if (op.has_ws) y.K2 = op.in, y.wt2 = P(o.ws_f), y.ldw2 = pad4(op.out);
                else y.flags |= F_ADD_SRC2;
                y.flags |= F_RELU | (grad ? F_SAVE_MASK : 0), y.mask_off = o.mask_y;
                if (last && lossprog) return fail(LINNA_EINVAL, "training needs a LINEAR last layer without activation");
                y.epi = last ? EPI_HEAD : EPI_ACT;
                y.dst = other(cur);
                if (train && !last) y.rm_off = rm_act[i + 1].off, y.rm_ld = rm_act[i + 1].ld;
                if (last && vjp) {
                    y.flags |= F_COT | F_OUT_VEC;
                    if (train) y.rm_off = rm_gz[i].off, y.rm_ld = rm_gz[i].ld;
                }
                cur = y.dst;
            }
            if (last) {
                Step &s = pg.steps[ns - 1];
                if (pk == PROG_PREDICT) s.flags |= F_OUT_VEC;
                if (vjp && !(s.flags & F_COT)) s.flags |= F_COT | F_OUT_VEC;
                if (pk == PROG_GRAD && m->ypositive) s.flags |= F_SAVE_Y, s.ybuf = bufY;
            }
        }
        if (vjp) {
            // backward-data from the cotangent the head step left in `cur`, down to d out / d theta (EPI_GRAD with
            // input_theta); with the training buffers present every d out / d z is also kept row-major for the
            // weight-gradient kernel (parameter gradients of an arbitrary cotangent: autograd through model(x))
            for (int i = (int)m->ops.size() - 1; i >= 0; --i) {
                const OpHost &op = m->ops[i];
                const OpOffsets &o = off[i];
                int pmask = -1;
                if (i > 0) {
                    const OpHost &pv = m->ops[i - 1];
                    if (pv.kind == LINNA_OP_RES || pv.act == LINNA_ACT_RELU) pmask = off[i - 1].mask_y;
                }
                if (op.kind == LINNA_OP_LINEAR) {
                    Step &s2 = new_step();
                    s2.src1 = cur, s2.K1 = op.out, s2.wt1 = P(o.w_b), s2.ldw1 = pad4(op.in), s2.N = op.in;
                    s2.epi = i == 0 ? EPI_GRAD : EPI_BWD;
                    if (pmask >= 0) s2.flags |= F_APPLY_MASK, s2.mask_off = pmask;
                    if (train && i > 0) s2.rm_off = rm_gz[i - 1].off, s2.rm_ld = rm_gz[i - 1].ld;
                    s2.dst = other(cur);
                    cur = s2.dst;
                } else {
                    Step &h = new_step();
                    h.src1 = cur, h.K1 = op.out, h.wt1 = P(o.w2_b), h.ldw1 = pad4(op.mid), h.N = op.mid;
                    h.scale = op.alpha, h.epi = EPI_BWD, h.flags = F_APPLY_MASK, h.mask_off = o.mask_h, h.dst = bufH;
                    if (train) h.rm_off = rm_gzh[i].off, h.rm_ld = rm_gzh[i].ld;
                    Step &x = new_step();
                    if (op.has_ws) {
                        x.src1 = cur, x.K1 = op.out, x.wt1 = P(o.ws_b), x.ldw1 = pad4(op.in);
                        x.src2 = bufH, x.K2 = op.mid, x.wt2 = P(o.w_b), x.ldw2 = pad4(op.in);
                    } else {
                        x.src1 = bufH, x.K1 = op.mid, x.wt1 = P(o.w_b), x.ldw1 = pad4(op.in);
                        x.src2 = cur, x.flags |= F_ADD_SRC2;
                    }
                    x.N = op.in, x.epi = i == 0 ? EPI_GRAD : EPI_BWD;
                    if (pmask >= 0) x.flags |= F_APPLY_MASK, x.mask_off = pmask;
                    if (train && i > 0) x.rm_off = rm_gz[i - 1].off, x.rm_ld = rm_gz[i - 1].ld;
                    x.dst = other(cur);
                    cur = x.dst;
                }
            }
        } else if (lossprog) {
            // q = delta @ Chat^-1 ; chi2 = q . delta ; (training) g_yhat = -2 q mask / (cmd B)
            const int dbuf = cur;
            Step &q = new_step();
            q.src1 = dbuf, q.K1 = n_out, q.wt1 = P(o_icov), q.ldw1 = pad4(n_out), q.N = n_out, q.epi = EPI_LOSSQ;
            q.dst = other(dbuf);
            if (train) {
                q.flags |= F_LOSS_GRAD, q.mask_off = mask_loss;
                q.rm_off = rm_gz.back().off, q.rm_ld = rm_gz.back().ld;
                cur = q.dst;
                for (int i = (int)m->ops.size() - 1; i >= 0; --i) {
                    const OpHost &op = m->ops[i];
                    const OpOffsets &o = off[i];
                    int pmask = -1;
                    if (i > 0) {
                        const OpHost &pv = m->ops[i - 1];
                        if (pv.kind == LINNA_OP_RES || pv.act == LINNA_ACT_RELU) pmask = off[i - 1].mask_y;
                    }
                    if (op.kind == LINNA_OP_LINEAR) {
                        if (i == 0) break;   // d loss / d xhat is not needed for training
                        Step &s2 = new_step();
                        s2.src1 = cur, s2.K1 = op.out, s2.wt1 = P(o.w_b), s2.ldw1 = pad4(op.in), s2.N = op.in;
                        s2.epi = EPI_BWD;
                        if (pmask >= 0) s2.flags |= F_APPLY_MASK, s2.mask_off = pmask;
                        s2.rm_off = rm_gz[i - 1].off, s2.rm_ld = rm_gz[i - 1].ld;
                        s2.dst = other(cur);
                        cur = s2.dst;
                    } else {
                        Step &h = new_step();
                        h.src1 = cur, h.K1 = op.out, h.wt1 = P(o.w2_b), h.ldw1 = pad4(op.mid), h.N = op.mid;
                        h.scale = op.alpha, h.epi = EPI_BWD, h.flags = F_APPLY_MASK, h.mask_off = o.mask_h, h.dst = bufH;
                        h.rm_off = rm_gzh[i].off, h.rm_ld = rm_gzh[i].ld;
                        if (i == 0) break;
                        Step &x = new_step();
                        if (op.has_ws) {
                            x.src1 = cur, x.K1 = op.out, x.wt1 = P(o.ws_b), x.ldw1 = pad4(op.in);
                            x.src2 = bufH, x.K2 = op.mid, x.wt2 = P(o.w_b), x.ldw2 = pad4(op.in);
                        } else {
                            x.src1 = bufH, x.K1 = op.mid, x.wt1 = P(o.w_b), x.ldw1 = pad4(op.in);
                            x.src2 = cur, x.flags |= F_ADD_SRC2;
                        }
                        x.N = op.in, x.epi = EPI_BWD;
                        if (pmask >= 0) x.flags |= F_APPLY_MASK, x.mask_off = pmask;
                        x.rm_off = rm_gz[i - 1].off, x.rm_ld = rm_gz[i - 1].ld;
                        x.dst = other(cur);
                        cur = x.dst;
                    }
                }
            }
        } else if (folded) {
            const OpHost &lastop = m->ops.back();
            const int K = lastop.in, sbuf = cur;
            Step &q = new_step();
            q.src1 = sbuf, q.K1 = K, q.wt1 = P(o_foldF), q.ldw1 = pad4(n_out), q.N = n_out, q.bias = P(o_foldc);
            q.epi = EPI_CHI2, q.dst = other(sbuf);
            if (pk == PROG_GRAD) {
                q.flags |= F_STORE_DST;
                // d lnL / d s = -(1/T) Af^T r, masked by the relu of the layer that produced s
                const int nops = (int)m->ops.size();
                Step &g0 = new_step();
                g0.src1 = q.dst, g0.K1 = n_out, g0.wt1 = P(o_foldB), g0.ldw1 = pad4(K), g0.N = K, g0.epi = EPI_BWD;
                g0.colscale = P(o_foldcs);
                const OpHost &pv = m->ops[nops - 2];
                if (pv.kind == LINNA_OP_RES || pv.act == LINNA_ACT_RELU) g0.flags |= F_APPLY_MASK, g0.mask_off = off[nops - 2].mask_y;
                g0.dst = sbuf;
                cur = g0.dst;
                for (int i = nops - 2; i >= 0; --i) {
                    const OpHost &op = m->ops[i];
                    const OpOffsets &o = off[i];
                    int pmask = -1;
                    if (i > 0) {
                        const OpHost &pp = m->ops[i - 1];
                        if (pp.kind == LINNA_OP_RES || pp.act == LINNA_ACT_RELU) pmask = off[i - 1].mask_y;
                    }
                    if (op.kind == LINNA_OP_LINEAR) {
                        Step &s2 = new_step();
                        s2.src1 = cur, s2.K1 = op.out, s2.wt1 = P(o.w_b), s2.ldw1 = pad4(op.in), s2.N = op.in;
                        s2.epi = i == 0 ? EPI_GRAD : EPI_BWD;
                        if (pmask >= 0) s2.flags |= F_APPLY_MASK, s2.mask_off = pmask;
                        s2.dst = other(cur);
                        cur = s2.dst;
                    } else {
                        Step &h = new_step();
                        h.src1 = cur, h.K1 = op.out, h.wt1 = P(o.w2_b), h.ldw1 = pad4(op.mid), h.N = op.mid;
                        h.scale = op.alpha, h.epi = EPI_BWD, h.flags = F_APPLY_MASK, h.mask_off = o.mask_h, h.dst = bufH;
                        Step &x = new_step();
                        if (op.has_ws) {
                            x.src1 = cur, x.K1 = op.out, x.wt1 = P(o.ws_b), x.ldw1 = pad4(op.in);
                            x.src2 = bufH, x.K2 = op.mid, x.wt2 = P(o.w_b), x.ldw2 = pad4(op.in);
                        } else {
                            x.src1 = bufH, x.K1 = op.mid, x.wt1 = P(o.w_b), x.ldw1 = pad4(op.in);
                            x.src2 = cur, x.flags |= F_ADD_SRC2;
                        }
                        x.N = op.in;
                        x.epi = i == 0 ? EPI_GRAD : EPI_BWD;
                        if (pmask >= 0) x.flags |= F_APPLY_MASK, x.mask_off = pmask;
                        x.dst = other(cur);
                        cur = x.dst;
                    }
                }
            }
        } else if (pk != PROG_PREDICT) {
            const int dbuf = cur;
            Step &q = new_step();
            q.src1 = dbuf, q.K1 = n_out, q.wt1 = P(o_quadF), q.ldw1 = pad4(n_out), q.N = n_out, q.epi = EPI_CHI2;
            q.dst = other(dbuf);
            const bool chol = m->quad_kind == LINNA_QUAD_CHOL;
            if (pk == PROG_GRAD && chol) q.flags |= F_STORE_DST;
            if (pk == PROG_GRAD) {
                // g_yhat = -(1/T) * Q_b r  (.) sigma*y_std (.* y if ypositive)
                Step &g0 = new_step();
                g0.src1 = chol ? q.dst : dbuf;
                g0.K1 = n_out, g0.wt1 = P(o_quadB), g0.ldw1 = pad4(n_out), g0.N = n_out, g0.epi = EPI_BWD;
                g0.colscale = P(o_cs);
                if (m->ypositive) g0.flags |= F_MUL_YSAVE, g0.ybuf = bufY;
                const OpHost &lastop = m->ops.back();
                const bool last_relu = lastop.kind == LINNA_OP_RES || lastop.act == LINNA_ACT_RELU;
                if (last_relu) g0.flags |= F_APPLY_MASK, g0.mask_off = off.back().mask_y;
                g0.dst = chol ? dbuf : other(dbuf);
                cur = g0.dst;
                if (m->has_extra) {
                    Step &gx = new_step();
                    gx.src1 = cur, gx.K1 = n_out, gx.wt1 = P(o_extra_b), gx.ldw1 = pad4(n_in), gx.N = n_in;
                    gx.epi = EPI_BWD, gx.dst = bufGX;
                }
                for (int i = (int)m->ops.size() - 1; i >= 0; --i) {
                    const OpHost &op = m->ops[i];
                    const OpOffsets &o = off[i];
                    // mask of the producer of this op's input
                    int pmask = -1;
                    if (i > 0) {
                        const OpHost &pv = m->ops[i - 1];
                        if (pv.kind == LINNA_OP_RES || pv.act == LINNA_ACT_RELU) pmask = off[i - 1].mask_y;
                    }
                    if (op.kind == LINNA_OP_LINEAR) {
                        Step &s = new_step();
                        s.src1 = cur, s.K1 = op.out, s.wt1 = P(o.w_b), s.ldw1 = pad4(op.in), s.N = op.in;
                        s.epi = i == 0 ? EPI_GRAD : EPI_BWD;
                        if (pmask >= 0) s.flags |= F_APPLY_MASK, s.mask_off = pmask;
                        if (i == 0 && m->has_extra) s.flags |= F_ADD_SRC2, s.src2 = bufGX;
                        s.dst = other(cur);
                        cur = s.dst;
                    } else {
                        if (i == 0 && m->has_extra) return fail(LINNA_EINVAL, "extra linear branch needs a LINEAR first op");
                        Step &h = new_step();
                        h.src1 = cur, h.K1 = op.out, h.wt1 = P(o.w2_b), h.ldw1 = pad4(op.mid), h.N = op.mid;
                        h.scale = op.alpha, h.epi = EPI_BWD, h.flags = F_APPLY_MASK, h.mask_off = o.mask_h, h.dst = bufH;
                        Step &x = new_step();
                        if (op.has_ws) {
                            x.src1 = cur, x.K1 = op.out, x.wt1 = P(o.ws_b), x.ldw1 = pad4(op.in);
                            x.src2 = bufH, x.K2 = op.mid, x.wt2 = P(o.w_b), x.ldw2 = pad4(op.in);
                        } else {
                            x.src1 = bufH, x.K1 = op.mid, x.wt1 = P(o.w_b), x.ldw1 = pad4(op.in);
                            x.src2 = cur, x.flags |= F_ADD_SRC2;
                        }
                        x.N = op.in;
                        x.epi = i == 0 ? EPI_GRAD : EPI_BWD;
                        if (pmask >= 0) x.flags |= F_APPLY_MASK, x.mask_off = pmask;
                        x.dst = other(cur);
                        cur = x.dst;
                    }
                }
            }
        }
        pg.n_steps = ns;
        m->prog_valid[pk] = true;
    }
    if (!m->prog_dev) CUDA_TRY(cudaMalloc(&m->prog_dev, sizeof(Program) * PROG_COUNT));
    CUDA_TRY(cudaMemcpy(m->prog_dev, m->prog_host, sizeof(Program) * PROG_COUNT, cudaMemcpyHostToDevice));

    // ---- scratch arenas sized for the largest grid
    for (int i = 0; i < 3; ++i)
        if (!m->occ[i]) m->occ[i] = fused_ffma_max_ctas_per_sm(1 << i);
    int max_ctas = m->num_sms * std::max(m->occ[0], std::max(m->occ[1], m->occ[2]));
    size_t need_arena = (size_t)max_ctas * arena_features * 32 * sizeof(float);
    size_t need_masks = (size_t)max_ctas * std::max(mask_total, 1) * 4;
    if (need_arena > m->arena_bytes) {
        if (m->arena) cudaFree(m->arena);
        m->arena = nullptr;
        CUDA_TRY(cudaMalloc(&m->arena, need_arena));
        m->arena_bytes = need_arena;
    }
    if (need_masks > m->masks_bytes) {
        if (m->masks) cudaFree(m->masks);
        m->masks = nullptr;
        CUDA_TRY(cudaMalloc(&m->masks, need_masks));
        m->masks_bytes = need_masks;
    }
    if (m->has_train) {
        // weight-gradient layer table and tile list (one launch covers every linear map)
        std::vector<WgradLayer> layers;
        auto add_layer = [&](const RM &gz, const RM &x, int N, int K, int wf, int bf, float gs) {
            WgradLayer L;
            memset(&L, 0, sizeof L);
            L.gz_off = gz.off, L.gz_ld = gz.ld, L.x_off = x.off, L.x_ld = x.ld, L.N = N, L.K = K;
            L.w_flat = wf, L.b_flat = bf, L.gscale = gs;
            layers.push_back(L);
        };
        for (size_t i = 0; i < m->ops.size(); ++i) {
            const OpHost &op = m->ops[i];
            if (op.kind == LINNA_OP_LINEAR) {
                add_layer(rm_gz[i], rm_act[i], op.out, op.in, fo[i].w, fo[i].b, 1.f);
            } else {
                add_layer(rm_gzh[i], rm_act[i], op.mid, op.in, fo[i].w, fo[i].b, 1.f);
                add_layer(rm_gz[i], rm_hid[i], op.out, op.mid, fo[i].w2, fo[i].b2, op.alpha);
                if (op.has_ws) add_layer(rm_gz[i], rm_act[i], op.out, op.in, fo[i].ws, -1, 1.f);
            }
        }
        std::vector<WgradTile> tiles;
        for (size_t l = 0; l < layers.size(); ++l) {
            const int kmax = layers[l].K + (layers[l].b_flat >= 0 ? 1 : 0);
            for (int n0 = 0; n0 < layers[l].N; n0 += 64)
                for (int k0 = 0; k0 < kmax; k0 += 64) tiles.push_back(WgradTile{(int32_t)l, n0, k0, 0});
        }
        if (m->wg_layers_dev) cudaFree(m->wg_layers_dev);
        if (m->wg_tiles_dev) cudaFree(m->wg_tiles_dev);
        if (m->map_fwd_dev) cudaFree(m->map_fwd_dev);
        if (m->map_bwd_dev) cudaFree(m->map_bwd_dev);
        m->wg_layers_dev = nullptr, m->wg_tiles_dev = nullptr, m->map_fwd_dev = nullptr, m->map_bwd_dev = nullptr;
        CUDA_TRY(cudaMalloc(&m->wg_layers_dev, layers.size() * sizeof(WgradLayer)));
        CUDA_TRY(cudaMalloc(&m->wg_tiles_dev, tiles.size() * sizeof(WgradTile)));
        CUDA_TRY(cudaMalloc(&m->map_fwd_dev, nflat * sizeof(int32_t)));
        CUDA_TRY(cudaMalloc(&m->map_bwd_dev, nflat * sizeof(int32_t)));
        CUDA_TRY(cudaMemcpy(m->wg_layers_dev, layers.data(), layers.size() * sizeof(WgradLayer), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(m->wg_tiles_dev, tiles.data(), tiles.size() * sizeof(WgradTile), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(m->map_fwd_dev, map_fwd.data(), nflat * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(m->map_bwd_dev, map_bwd.data(), nflat * sizeof(int32_t), cudaMemcpyHostToDevice));
        m->n_wg_tiles = (int)tiles.size();
        if (rm_total > m->rm_floats) {
            if (m->rm) cudaFree(m->rm);
            m->rm = nullptr;
            CUDA_TRY(cudaMalloc(&m->rm, rm_total * sizeof(float)));
            m->rm_floats = rm_total;
        }
        CUDA_TRY(cudaMemset(m->rm, 0, m->rm_floats * sizeof(float)));
        for (size_t i = 0; i < m->ops.size(); ++i) {   // the bias column of every layer input is 1
            const OpHost &op = m->ops[i];
            CUDA_TRY(launch_fill_col(m->rm + rm_act[i].off, rm_act[i].ld, op.in, m->max_batch, 1.f, 0));
            if (op.kind == LINNA_OP_RES)
                CUDA_TRY(launch_fill_col(m->rm + rm_hid[i].off, rm_hid[i].ld, op.mid, m->max_batch, 1.f, 0));
        }
    }
    CUDA_TRY(cudaDeviceSynchronize());
    if (m->has_train) {
        // tensor-core training kernels (tg_gemm.cu); the FP32 kernels above stay as the path for shapes they do not cover
        std::vector<std::array<int, 5>> flat_off(m->ops.size());
        std::vector<std::array<const float *, 2>> bias_ptr(m->ops.size());
        std::vector<std::array<int64_t, 8>> blob_off(m->ops.size());
        for (size_t i = 0; i < m->ops.size(); ++i) {
            const bool res = m->ops[i].kind == LINNA_OP_RES;
            flat_off[i] = {fo[i].w, fo[i].b, fo[i].w2, fo[i].b2, fo[i].ws};
            bias_ptr[i] = {P(off[i].b), res ? P(off[i].b2) : nullptr};
            blob_off[i] = {(int64_t)off[i].w_f, (int64_t)off[i].w_b, (int64_t)off[i].b, res ? (int64_t)off[i].w2_f : -1,
                           res ? (int64_t)off[i].w2_b : -1, res ? (int64_t)off[i].b2 : -1,
                           res && m->ops[i].has_ws ? (int64_t)off[i].ws_f : -1, res && m->ops[i].has_ws ? (int64_t)off[i].ws_b : -1};
        }
        m->tg_why.clear();
        if (!getenv("LINNA_TRAIN_NO_TC")) m->tg = tg_build(m, flat_off, bias_ptr, blob_off, m->tg_why);
        else m->tg_why = "LINNA_TRAIN_NO_TC set";
        if (!m->tg)
            fprintf(stderr, "linna_b200: tensor-core training kernels unavailable for this model (%s); training runs on the FP32 "
                            "FFMA kernels\n", m->tg_why.c_str());
        cudaGetLastError();
    }
    return LINNA_OK;
}

static int copy_ops(linna_model *m, const linna_op_desc_t *ops, int n_ops, bool check_shapes)
{
    if (check_shapes && (int)m->ops.size() != n_ops) return fail(LINNA_EINVAL, "op count changed");
    std::vector<OpHost> v(n_ops);
    int64_t np = 0;
    int width = m->n_in;
    for (int i = 0; i < n_ops; ++i) {
        const linna_op_desc_t &d = ops[i];
        OpHost &o = v[i];
        o.kind = d.kind, o.in = d.in_dim, o.mid = d.mid_dim, o.out = d.out_dim, o.act = d.act, o.alpha = d.alpha;
        if (d.in_dim != width) return fail(LINNA_EINVAL, "op %d: in_dim %d does not chain (expected %d)", i, d.in_dim, width);
        if (d.in_dim <= 0 || d.out_dim <= 0) return fail(LINNA_EINVAL, "op %d: bad dims", i);
        if (!d.w || !d.b) return fail(LINNA_EINVAL, "op %d: null weights", i);
        if (d.kind == LINNA_OP_LINEAR) {
            o.w.assign(d.w, d.w + (size_t)d.out_dim * d.in_dim);
            o.b.assign(d.b, d.b + d.out_dim);
            np += (int64_t)d.out_dim * d.in_dim + d.out_dim;
        } else if (d.kind == LINNA_OP_RES) {
            if (d.mid_dim <= 0 || !d.w2 || !d.b2) return fail(LINNA_EINVAL, "op %d: bad res block", i);
            if (!d.ws && d.in_dim != d.out_dim) return fail(LINNA_EINVAL, "op %d: identity skip needs in == out", i);
            o.w.assign(d.w, d.w + (size_t)d.mid_dim * d.in_dim);
            o.b.assign(d.b, d.b + d.mid_dim);
            o.w2.assign(d.w2, d.w2 + (size_t)d.out_dim * d.mid_dim);
            o.b2.assign(d.b2, d.b2 + d.out_dim);
            o.has_ws = d.ws != nullptr;
            if (o.has_ws) o.ws.assign(d.ws, d.ws + (size_t)d.out_dim * d.in_dim);
            np += (int64_t)d.mid_dim * d.in_dim + d.mid_dim + (int64_t)d.out_dim * d.mid_dim + d.out_dim +
                  (o.has_ws ? (int64_t)d.out_dim * d.in_dim : 0);
        } else
            return fail(LINNA_EINVAL, "op %d: unknown kind %d", i, d.kind);
        if (check_shapes) {
            const OpHost &p = m->ops[i];
            if (p.kind != o.kind || p.in != o.in || p.mid != o.mid || p.out != o.out)
                return fail(LINNA_EINVAL, "op %d: shape changed", i);
        }
        width = d.out_dim;
    }
    if (width != m->n_out) return fail(LINNA_EINVAL, "last op out_dim %d != n_out %d", width, m->n_out);
    m->ops.swap(v);
    m->n_params = np;
    return LINNA_OK;
}

// ----------------------------------------------------------------------------------------------
extern "C" {

int linna_abi_version(void) { return LINNA_ABI_VERSION; }
const char *linna_last_error(void) { return g_err.c_str(); }
int64_t linna_launch_count(void) { return g_launches.load(); }

int linna_model_create(const linna_model_desc_t *d, int device, linna_model_t **out)
{
    if (!d || !out) return fail(LINNA_EINVAL, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LINNA_ENODEV, "no CUDA device: linna_b200 has no CPU path");
    if (device < 0 || device >= ndev) return fail(LINNA_ENODEV, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(LINNA_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    if (d->n_in <= 0 || d->n_out <= 0 || d->n_ops <= 0 || !d->ops) return fail(LINNA_EINVAL, "bad model dims");
    if (!d->x_mean || !d->x_std || !d->y_mean || !d->y_std) return fail(LINNA_EINVAL, "null transform arrays");
    linna_model *m = new linna_model();
    m->device = device, m->num_sms = prop.multiProcessorCount;
    m->n_in = d->n_in, m->n_out = d->n_out, m->ypositive = d->ypositive ? 1 : 0;
    int rc = copy_ops(m, d->ops, d->n_ops, false);
    if (rc) { delete m; return rc; }
    m->x_mean.assign(d->x_mean, d->x_mean + d->n_in);
    m->x_std.assign(d->x_std, d->x_std + d->n_in);
    m->y_mean.assign(d->y_mean, d->y_mean + d->n_out);
    m->y_std.assign(d->y_std, d->y_std + d->n_out);
    if (d->sigma) m->sigma.assign(d->sigma, d->sigma + d->n_out);
    else m->sigma.assign(d->n_out, 1.f);
    m->log10_flag.assign(d->n_in, 0);
    if (d->log10_flag)
        for (int i = 0; i < d->n_in; ++i) { m->log10_flag[i] = d->log10_flag[i] ? 1 : 0; m->has_log10 |= d->log10_flag[i] != 0; }
    if (d->extra_linear_w) {
        if (!d->extra_linear_b) { delete m; return fail(LINNA_EINVAL, "extra_linear_b is null"); }
        m->has_extra = true;
        m->extra_w.assign(d->extra_linear_w, d->extra_linear_w + (size_t)d->n_out * d->n_in);
        m->extra_b.assign(d->extra_linear_b, d->extra_linear_b + d->n_out);
        m->extra_scale = d->extra_linear_scale;
        m->n_params += (int64_t)d->n_out * d->n_in + d->n_out;
    }
    DeviceGuard guard(device);
    rc = rebuild(m);
    if (rc) { free_device(m); delete m; return rc; }
    if (cudaStreamCreateWithFlags(&m->hstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->cstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->dstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->last_done, cudaEventDisableTiming) != cudaSuccess) {
        free_device(m); delete m;
        return fail(LINNA_ECUDA, "cudaStreamCreate/cudaEventCreate failed");
    }
    *out = m;
    return LINNA_OK;
}

void linna_model_destroy(linna_model_t *m)
{
    if (!m) return;
    DeviceGuard guard(m->device);
    cudaDeviceSynchronize();
    if (m->helper) { delete m->helper; m->helper = nullptr; }
    if (m->hstream) cudaStreamDestroy(m->hstream);
    if (m->cstream) cudaStreamDestroy(m->cstream);
    if (m->dstream) cudaStreamDestroy(m->dstream);
    for (cudaEvent_t e : m->pipe_events) cudaEventDestroy(e);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    if (m->h_in_stage) cudaFreeHost(m->h_in_stage);
    if (m->last_done) cudaEventDestroy(m->last_done);
    if (m->tc) tc_destroy(m->tc);
    if (m->d_in) cudaFree(m->d_in);
    if (m->d_out) cudaFree(m->d_out);
    if (m->d_lnp) cudaFree(m->d_lnp);
    if (m->d_grad) cudaFree(m->d_grad);
    free_device(m);
    delete m;
}

int linna_model_set_likelihood(linna_model_t *m, const linna_like_desc_t *l)
{
    if (!m || !l) return fail(LINNA_EINVAL, "null argument");
    if (!l->prior_kind || !l->prior_arg1 || !l->prior_arg2 || !l->data || !l->quad)
        return fail(LINNA_EINVAL, "null likelihood arrays");
    if (!(l->temperature > 0)) return fail(LINNA_EINVAL, "temperature must be > 0");
    if (l->quad_kind != LINNA_QUAD_CHOL && l->quad_kind != LINNA_QUAD_DENSE) return fail(LINNA_EINVAL, "bad quad_kind");
    m->prior_kind.assign(l->prior_kind, l->prior_kind + m->n_in);
    m->prior_scale.resize(m->n_in), m->prior_shift.resize(m->n_in);
    for (int i = 0; i < m->n_in; ++i) {
        if (l->prior_kind[i] == LINNA_PRIOR_GAUSS) m->prior_scale[i] = l->prior_arg2[i];
        else if (l->prior_kind[i] == LINNA_PRIOR_FLAT) m->prior_scale[i] = l->prior_arg2[i] - l->prior_arg1[i];
        else return fail(LINNA_EINVAL, "prior %d: unknown dist %d", i, l->prior_kind[i]);  // main.py:128-129
        m->prior_shift[i] = l->prior_arg1[i];
    }
    m->data.assign(l->data, l->data + m->n_out);
    m->quad.assign(l->quad, l->quad + (size_t)m->n_out * m->n_out);
    m->quad_kind = l->quad_kind;
    if (l->quad_kind == LINNA_QUAD_DENSE) {  // x^T A x == x^T sym(A) x; the gradient then is -sym(A) d / T
        const int n = m->n_out;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) {
                float s = 0.5f * (m->quad[(size_t)i * n + j] + m->quad[(size_t)j * n + i]);
                m->quad[(size_t)i * n + j] = m->quad[(size_t)j * n + i] = s;
            }
    }
    m->temperature = l->temperature;
    m->has_like = true;
    return rebuild(m);
}

int linna_model_set_weights(linna_model_t *m, const linna_op_desc_t *ops, int32_t n_ops, const float *ew, const float *eb)
{
    if (!m || !ops) return fail(LINNA_EINVAL, "null argument");
    int rc = copy_ops(m, ops, n_ops, true);
    if (rc) return rc;
    if (m->has_extra) {
        if (!ew || !eb) return fail(LINNA_EINVAL, "extra linear weights required");
        m->extra_w.assign(ew, ew + (size_t)m->n_out * m->n_in);
        m->extra_b.assign(eb, eb + m->n_out);
        m->n_params += (int64_t)m->n_out * m->n_in + m->n_out;
    }
    return rebuild(m);
}

int linna_model_info(const linna_model_t *m, int32_t *n_in, int32_t *n_out, int64_t *n_params, int32_t *num_sms)
{
    if (!m) return fail(LINNA_EINVAL, "null model");
    if (n_in) *n_in = m->n_in;
    if (n_out) *n_out = m->n_out;
    if (n_params) *n_params = m->n_params;
    if (num_sms) *num_sms = m->num_sms;
    return LINNA_OK;
}

int linna_model_set_path(linna_model_t *m, int32_t path, int64_t tc_min_rows)
{
    if (!m || path < 0 || path > 3) return fail(LINNA_EINVAL, "path must be 0 (auto), 1 (FFMA), 2 (tensor core) or 3 (cluster)");
    m->path = path;
    if (tc_min_rows > 0) m->tc_min_rows = tc_min_rows;
    return LINNA_OK;
}

int linna_model_last_kernel(const linna_model_t *m) { return m ? m->last_kernel : 0; }

int linna_debug_tc_counters(linna_model_t *m, int64_t *out, int32_t max_ctas)
{
    if (!m || !out) return 0;
    return tc_debug_read(m->tc, reinterpret_cast<long long *>(out), max_ctas);
}

int linna_debug_cluster_counters(int64_t *out, int32_t max_values)
{
    if (!out) return 0;
    return cluster_ffma_debug_read(reinterpret_cast<long long *>(out), max_values);
}

int linna_debug_tg_counters(linna_model_t *m, int64_t *out, int32_t max_steps)
{
    if (!m || !out) return 0;
    return tg_debug_read(m->tg, reinterpret_cast<long long *>(out), max_steps);
}

int linna_model_set_fold(linna_model_t *m, int32_t on)
{
    if (!m) return fail(LINNA_EINVAL, "null model");
    m->fold_enabled = on != 0;
    return rebuild(m);
}

int linna_model_set_tile_rows(linna_model_t *m, int32_t rows)
{
    if (!m || (rows != 0 && rows != 8 && rows != 16 && rows != 32)) return fail(LINNA_EINVAL, "rows must be 0, 8, 16 or 32");
    m->force_rows = rows;
    return LINNA_OK;
}

// Launch geometry of the small-batch cluster kernel for program pk: the largest cluster the device schedules (16 CTAs,
// else 8) with the deepest weight ring that fits next to the shared-memory activation arena.
static bool cluster_resolve(linna_model *m, int pk)
{
    if (m->cl_state[pk]) return m->cl_state[pk] > 0;
    m->cl_state[pk] = -1;
    const Program &pg = m->prog_host[pk];
    int maxN = 0;
    for (int i = 0; i < pg.n_steps; ++i) {
        maxN = std::max(maxN, pg.steps[i].N);
        if (pg.steps[i].rm_off >= 0 || (pg.steps[i].flags & F_COT)) { m->cl_why = "program keeps row-major copies"; return false; }
    }
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device);
    const char *ecs = getenv("LINNA_CLUSTER_SIZE");
    const int want = ecs ? atoi(ecs) : 16;
    for (int cs : {16, 8}) {
        if (cs > want) continue;
        if (maxN > 256 * cs) continue;
        for (int depth : {16, 8}) {
            int chi_q = 0;
            const size_t smem = cluster_ffma_smem_bytes(pg, m->n_out, cs, depth, &chi_q);
            if (smem + 8192 > (size_t)max_smem) continue;   // the kernel's static shared memory (step table) comes on top
            const int nc = cluster_ffma_max_clusters(cs, depth, smem);
            if (nc < 1) continue;
            m->cl_cs[pk] = cs, m->cl_depth[pk] = depth, m->cl_clusters[pk] = nc, m->cl_chi_q[pk] = chi_q, m->cl_smem[pk] = smem;
            m->cl_state[pk] = 1;
            return true;
        }
    }
    m->cl_why = "activation arena does not fit in shared memory / no cluster can be scheduled";
    return false;
}

// lnP must not depend on whether the gradient was asked for: the likelihood and the gradient program take the cluster
// kernel together or not at all (the gradient program carries the relu masks and may not fit where the other does).
static bool cluster_usable(linna_model *m, int pk)
{
    if (pk == PROG_LNP || pk == PROG_GRAD) {
        for (int q : {PROG_LNP, PROG_GRAD})
            if (m->prog_valid[q] && !cluster_resolve(m, q)) {
                m->cl_state[pk] = -1;
                return false;
            }
    }
    return cluster_resolve(m, pk);
}

static int run(linna_model *m, int pk, const float *in, int64_t n, float *out_vec, int out_kind, float *lnp, float *grad,
               int input_theta, cudaStream_t stream, const KernelArgs *proto = nullptr)
{
    if (!m) return fail(LINNA_EINVAL, "null model");
    if (n < 0) return fail(LINNA_EINVAL, "negative n");
    if (n == 0) return LINNA_OK;
    if (!in) return fail(LINNA_EINVAL, "null input");
    if (!m->prog_valid[pk]) return fail(LINNA_ESTATE, "likelihood constants not set (linna_model_set_likelihood)");
    LINNA_ON_DEVICE(m);
    if (m->have_last && m->last_stream != stream) CUDA_TRY(cudaStreamWaitEvent(stream, m->last_done, 0));
    // Small batches (an emcee ensemble of a few walkers, a handful of HMC chains, one predict call): the cluster kernel,
    // which spreads every layer over the CTAs of a thread-block cluster instead of streaming all weights through one SM.
    if ((pk == PROG_PREDICT || pk == PROG_LNP || pk == PROG_GRAD) && !proto &&
        (m->path == 3 || (m->path == 0 && !m->force_rows && n <= m->cl_max_rows && n < m->tc_min_rows))) {
        // auto mode: up to two tiles of 8 walkers per schedulable cluster (beyond that the tiled kernel, which puts one
        // tile on every SM, is faster: scratch/cluster_time.py)
        if (cluster_usable(m, pk) && (m->path == 3 || n <= 16 * (int64_t)m->cl_clusters[pk])) {
            KernelArgs a;
            memset(&a, 0, sizeof a);
            a.prog = m->prog_dev + pk, a.c = m->consts, a.in = in, a.out_vec = out_vec, a.lnp = lnp, a.grad = grad;
            a.n = n, a.input_theta = input_theta, a.out_kind = out_kind;
            const int clusters = (int)std::min<int64_t>((n + 7) / 8, m->cl_clusters[pk]);
            CUDA_TRY(launch_cluster_ffma(a, m->cl_chi_q[pk], m->cl_cs[pk], m->cl_depth[pk], clusters, m->cl_smem[pk], stream));
            g_launches.fetch_add(1);
            m->last_kernel = 3;
            CUDA_TRY(cudaEventRecord(m->last_done, stream));
            m->last_stream = stream, m->have_last = true;
            return LINNA_OK;
        }
        if (m->path == 3 && m->cl_state[pk] < 0) return fail(LINNA_EINVAL, "cluster kernel unavailable: %s", m->cl_why.c_str());
    }
    // Large lnP / lnP+gradient batches go to the tensor-core (tcgen05) kernel; everything else stays on the
    // FP32 FFMA kernel.
    const bool tc_prog = pk == PROG_LNP || pk == PROG_GRAD || pk == PROG_PREDICT;
    if (tc_prog && !proto && m->path == 2 && (m->tc_failed || m->has_extra))
        return fail(LINNA_EINVAL, "tensor-core path unavailable: %s",
                    m->has_extra ? "extra linear branch not supported on the tensor-core path" : m->tc_why.c_str());
    if (tc_prog && !proto && m->path != 1 && m->path != 3 && !m->has_extra && !m->tc_failed && (m->path == 2 || n >= m->tc_min_rows)) {
        if (!m->tc) {
            std::string why;
            m->tc = tc_build(m, why);
            if (!m->tc) {
                m->tc_failed = true;
                m->tc_why = why;
                if (m->path == 2) return fail(LINNA_EINVAL, "tensor-core path unavailable: %s", why.c_str());
                // automatic selection: say (once per model build) that this model runs ~9x slower than it could
                fprintf(stderr, "linna_b200: tensor-core likelihood kernel unavailable for this model (%s); large batches run on the "
                                "FP32 FFMA kernel\n", why.c_str());
            }
        }
        const bool have = m->tc && (pk == PROG_PREDICT ? tc_has_predict(m->tc) : pk == PROG_GRAD ? tc_has_grad(m->tc) : tc_has_lnp(m->tc));
        if (m->tc && !have) {
            if (m->path == 2)
                return fail(LINNA_EINVAL, pk == PROG_GRAD ? "tensor-core gradient path needs the folded likelihood tail"
                                                          : "tensor-core path needs the Cholesky form of the quadratic");
        } else if (m->tc) {
            if (pk == PROG_LNP) CUDA_TRY(tc_launch_lnp(m, m->tc, in, n, lnp, stream));
            else if (pk == PROG_GRAD) CUDA_TRY(tc_launch_grad(m, m->tc, in, n, lnp, grad, stream));
            else CUDA_TRY(tc_launch_predict(m, m->tc, in, n, out_vec, out_kind, stream));
            g_launches.fetch_add(1);
            {
                // Fix-up: rows that came out NaN on the tensor-core path (an activation beyond the fp16 range) are redone
                // by the FP32 kernel, which has the reference's own range.  The row list and its length stay on the
                // device; with nothing flagged the launch is a few idle CTAs.
                KernelArgs a;
                memset(&a, 0, sizeof a);
                a.prog = m->prog_dev + pk, a.c = m->consts, a.in = in, a.lnp = lnp, a.grad = grad;
                a.out_vec = out_vec, a.out_kind = out_kind, a.input_theta = input_theta;
                a.arena = m->arena, a.masks = m->masks, a.n = pk == PROG_PREDICT ? 2 * n + 1024 : n;
                tc_fix_buffers(m->tc, &a.row_index, &a.n_dev, &a.zero_me);
                const int grid = (int)std::min<int64_t>((n + 7) / 8, std::min(m->num_sms, 32));
                CUDA_TRY(launch_fused_ffma(a, 1, grid, stream));
                tc_launch_done(m->tc);
                g_launches.fetch_add(1);
            }
            m->last_kernel = 2;
            CUDA_TRY(cudaEventRecord(m->last_done, stream));
            m->last_stream = stream, m->have_last = true;
            return LINNA_OK;
        }
    }
    // Launch plan: full waves of 32-row tiles, then the remainder in 16- and 8-row tiles, so that the
    // last wave costs a fraction of a full one instead of leaving most SMs idle for a whole tile time.
    int64_t done = 0;
    while (done < n) {
        const int64_t left = n - done;
        int rows = m->force_rows;
        int64_t take = left;
        if (!rows) {
            const int64_t g32 = (int64_t)m->num_sms * m->occ[2] * 32, g16 = (int64_t)m->num_sms * m->occ[1] * 16;
            if (left >= g32) rows = 32, take = left / g32 * g32;
            else if (left >= g16) rows = 16, take = left / g16 * g16;
            else if (left > (int64_t)m->num_sms * 8 * m->occ[0]) rows = 16;
            else rows = 8;   // every tile resident at once: the deep-ring 8-row kernel (one CTA per SM)
        }
        const int rg = rows / 8;
        const int occ = m->occ[rg == 4 ? 2 : rg == 2 ? 1 : 0];
        const int64_t tiles = (take + rows - 1) / rows;
        const int grid = (int)std::min<int64_t>(tiles, (int64_t)m->num_sms * occ);
        KernelArgs a;
        memset(&a, 0, sizeof a);
        a.prog = m->prog_dev + pk;
        a.c = m->consts;
        a.in = in + done * m->n_in;
        a.out_vec = out_vec ? out_vec + done * m->n_out : nullptr;
        a.lnp = lnp ? lnp + done : nullptr;
        a.grad = grad ? grad + done * m->n_in : nullptr;
        a.arena = m->arena, a.masks = m->masks;
        a.n = take, a.input_theta = input_theta, a.out_kind = out_kind;
        if (proto) {
            a.target = proto->target ? proto->target + done * m->n_out : nullptr;
            a.cmd = proto->cmd ? proto->cmd + done : nullptr;
            a.rm_base = proto->rm_base, a.rm_row0 = done;
            a.loss_inv_B = proto->loss_inv_B, a.delta_kind = proto->delta_kind;
        }
        CUDA_TRY(launch_fused_ffma(a, rg, grid, stream));
        g_launches.fetch_add(1);
        m->last_kernel = 1;
        done += take;
    }
    CUDA_TRY(cudaEventRecord(m->last_done, stream));
    m->last_stream = stream, m->have_last = true;
    return LINNA_OK;
}

int linna_predict(linna_model_t *m, const float *theta, int64_t n, float *out, int32_t out_kind, void *stream)
{
    if (out_kind < LINNA_OUT_YHAT || out_kind > LINNA_OUT_M) return fail(LINNA_EINVAL, "bad out_kind");
    if (n > 0 && !out) return fail(LINNA_EINVAL, "null output");
    return run(m, PROG_PREDICT, theta, n, out, out_kind, nullptr, nullptr, 1, (cudaStream_t)stream);
}

int linna_lnp(linna_model_t *m, const float *u, int64_t n, float *lnp, void *stream)
{
    if (n > 0 && !lnp) return fail(LINNA_EINVAL, "null output");
    return run(m, PROG_LNP, u, n, nullptr, 0, lnp, nullptr, 0, (cudaStream_t)stream);
}

int linna_lnp_grad(linna_model_t *m, const float *u, int64_t n, float *lnp, float *grad, void *stream)
{
    if (n > 0 && (!lnp || !grad)) return fail(LINNA_EINVAL, "null output");
    return run(m, PROG_GRAD, u, n, nullptr, 0, lnp, grad, 0, (cudaStream_t)stream);
}

static int ensure(float **p, size_t *cap, size_t need);

int linna_predict_vjp(linna_model_t *m, const float *theta, int64_t n, const float *cot, int32_t out_kind, float *out, float *gtheta,
                      float *gparams, void *stream)
{
    if (!m) return fail(LINNA_EINVAL, "null model");
    if (out_kind < LINNA_OUT_YHAT || out_kind > LINNA_OUT_M) return fail(LINNA_EINVAL, "bad out_kind");
    if (n < 0) return fail(LINNA_EINVAL, "negative n");
    if (n == 0) return LINNA_OK;
    if (!theta || !cot) return fail(LINNA_EINVAL, "null buffer");
    if (!m->prog_valid[PROG_VJP]) return fail(LINNA_EINVAL, "vector-Jacobian program unavailable for this model (extra linear branch)");
    if (gparams && (!m->has_train || n > m->max_batch))
        return fail(LINNA_ESTATE, "parameter gradients need linna_train_setup with max_batch >= n (%lld)", (long long)n);
    LINNA_ON_DEVICE(m);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (!gtheta) {
        if ((rc = ensure(&m->d_grad, &m->d_grad_cap, (size_t)n * m->n_in))) return rc;
        gtheta = m->d_grad;
    }
    KernelArgs p;
    memset(&p, 0, sizeof p);
    p.target = cot, p.rm_base = gparams ? m->rm : nullptr, p.loss_inv_B = 1.f;
    if ((rc = run(m, PROG_VJP, theta, n, out, out_kind, nullptr, gtheta, 1, st, &p))) return rc;
    if (gparams) {
        AdamArgs a;
        memset(&a, 0, sizeof a);
        a.grads = gparams, a.fuse = 0;
        CUDA_TRY(launch_wgrad(m->wg_layers_dev, m->wg_tiles_dev, m->n_wg_tiles, m->rm, (int)n, a, st));
        g_launches.fetch_add(1);
        CUDA_TRY(cudaEventRecord(m->last_done, st));
    }
    return LINNA_OK;
}

static int ensure(float **p, size_t *cap, size_t need)
{
    if (*cap >= need) return LINNA_OK;
    if (*p) cudaFree(*p);
    *p = nullptr, *cap = 0;
    size_t want = need + need / 4;
    if (cudaMalloc(p, want * sizeof(float)) != cudaSuccess) return fail(LINNA_ENOMEM, "cudaMalloc of %zu floats failed", want);
    *cap = want;
    return LINNA_OK;
}

int linna_predict_host(linna_model_t *m, const float *theta, int64_t n, float *out, int32_t out_kind)
{
    if (!m) return fail(LINNA_EINVAL, "null model");
    if (n <= 0) return n == 0 ? LINNA_OK : fail(LINNA_EINVAL, "negative n");
    if (!theta || !out) return fail(LINNA_EINVAL, "null buffer");
    LINNA_ON_DEVICE(m);
    int rc;
    if ((rc = ensure(&m->d_in, &m->d_in_cap, (size_t)n * m->n_in))) return rc;
    if ((rc = ensure(&m->d_out, &m->d_out_cap, (size_t)n * m->n_out))) return rc;
    CUDA_TRY(cudaMemcpyAsync(m->d_in, theta, (size_t)n * m->n_in * sizeof(float), cudaMemcpyHostToDevice, m->hstream));
    if ((rc = linna_predict(m, m->d_in, n, m->d_out, out_kind, m->hstream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, m->d_out, (size_t)n * m->n_out * sizeof(float), cudaMemcpyDeviceToHost, m->hstream));
    CUDA_TRY(cudaStreamSynchronize(m->hstream));
    return LINNA_OK;
}

// Host-buffer lnP / lnP+gradient: the batch is cut into chunks of one round of the tensor-core kernel
// (one walker pair of 256 rows on every CTA pair) and the three legs run on three streams, so that the
// host->device copy of chunk k+1 and the device->host copy of chunk k-1 hide behind the kernel of chunk k.
static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

static int ensure_pinned(float **p, size_t *cap, size_t need)
{
    if (*cap >= need) return LINNA_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr, *cap = 0;
    if (cudaHostAlloc(p, (need + need / 4) * sizeof(float), cudaHostAllocDefault) != cudaSuccess)
        return fail(LINNA_ENOMEM, "cudaHostAlloc of %zu floats failed", need);
    *cap = need + need / 4;
    return LINNA_OK;
}

static inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

// One pipelined host-buffer call, shared by the calling thread and the model's helper thread.
struct StageJob {
    // input staging: `npieces` pieces of `piece_rows` rows, claimed front to back through `next_piece` by whichever
    // thread is free; piece_done[q] is set (release) when piece q sits in the pinned staging buffer
    const float *u = nullptr;
    float *in_stage = nullptr;
    int64_t n = 0, piece_rows = 0;
    int n_in = 0, npieces = 0;
    std::atomic<int> next_piece{0};
    std::atomic<uint8_t> *piece_done = nullptr;
    // results: chunk k = rows [cstart[k], cstart[k+1]); `issued` counts the chunks whose device->host copy has been
    // enqueued (home[k] recorded); the helper hands them to the caller's arrays in order
    int nchunks = 0;
    const int64_t *cstart = nullptr;
    cudaEvent_t *events = nullptr;   // 3 per chunk, [3k + 2] = results of chunk k on the host
    std::atomic<int> issued{0};
    bool stage_in = false, stage_out = false;
    float *lnp = nullptr, *grad = nullptr;
    const float *s_lnp = nullptr, *s_grad = nullptr;
    std::atomic<int> abort{0}, helper_done{0}, helper_err{0};
};

static inline void stage_piece(StageJob *j, int q)
{
    const int64_t a0 = (int64_t)q * j->piece_rows, pr = std::min(j->piece_rows, j->n - a0);
    memcpy(j->in_stage + a0 * j->n_in, j->u + a0 * j->n_in, (size_t)pr * j->n_in * sizeof(float));
    j->piece_done[q].store(1, std::memory_order_release);
}

static inline void copy_out_chunk(StageJob *j, int k)
{
    const int64_t r0 = j->cstart[k], rows = j->cstart[k + 1] - r0;
    if (j->s_lnp != j->lnp) memcpy(j->lnp + r0, j->s_lnp + r0, (size_t)rows * sizeof(float));
    if (j->grad && j->s_grad != j->grad) memcpy(j->grad + r0 * j->n_in, j->s_grad + r0 * j->n_in, (size_t)rows * j->n_in * sizeof(float));
}

static void stage_helper_run(StageJob *j)
{
    if (j->stage_in)
        for (;;) {
            if (j->abort.load(std::memory_order_relaxed)) break;
            const int q = j->next_piece.fetch_add(1, std::memory_order_relaxed);
            if (q >= j->npieces) break;
            stage_piece(j, q);
        }
    if (j->stage_out)
        for (int k = 0; k < j->nchunks; ++k) {
            while (j->issued.load(std::memory_order_acquire) <= k && !j->abort.load(std::memory_order_relaxed)) cpu_relax();
            if (j->abort.load(std::memory_order_relaxed)) break;
            if (cudaEventSynchronize(j->events[3 * k + 2]) != cudaSuccess) {
                j->helper_err.store(1);
                break;
            }
            copy_out_chunk(j, k);
        }
    j->helper_done.store(1, std::memory_order_release);
}

static int lnp_host_pipelined(linna_model *m, const float *u, int64_t n, float *lnp, float *grad)
{
    LINNA_ON_DEVICE(m);
    int rc;
    const int n_in = m->n_in;
    if ((rc = ensure(&m->d_in, &m->d_in_cap, (size_t)n * n_in))) return rc;
    if ((rc = ensure(&m->d_lnp, &m->d_lnp_cap, (size_t)n))) return rc;
    if (grad && (rc = ensure(&m->d_grad, &m->d_grad_cap, (size_t)n * n_in))) return rc;
    // Pageable caller buffers (what an emcee / zeus caller holds: plain numpy arrays) go through pinned staging on both
    // sides.  cudaMemcpyAsync on pageable memory stages inside the driver on the CALLING thread (measured 17-20 GB/s:
    // 0.6-0.7 ms for the 12 MB of 10^5 C3 walkers, most of a kernel), and a device->host copy into pageable memory blocks
    // until the kernel before it has finished.  Here the calling thread and ONE helper thread fill the input staging
    // piece by piece ahead of the GPU, and the helper drains the result staging behind it.
    const size_t row_bytes = (size_t)n_in * sizeof(float);
    if ((size_t)n * row_bytes <= ((size_t)64 << 10)) {
        // Latency path (an emcee ensemble of a few walkers per call): nothing to pipeline -- one stream, pinned staging on
        // both sides, one synchronisation.
        const size_t out_floats = (size_t)n * (1 + (grad ? n_in : 0));
        if ((rc = ensure_pinned(&m->h_in_stage, &m->h_in_stage_cap, (size_t)n * n_in))) return rc;
        if ((rc = ensure_pinned(&m->h_stage, &m->h_stage_cap, out_floats))) return rc;
        memcpy(m->h_in_stage, u, (size_t)n * row_bytes);
        CUDA_TRY(cudaMemcpyAsync(m->d_in, m->h_in_stage, (size_t)n * row_bytes, cudaMemcpyHostToDevice, m->hstream));
        rc = grad ? linna_lnp_grad(m, m->d_in, n, m->d_lnp, m->d_grad, m->hstream) : linna_lnp(m, m->d_in, n, m->d_lnp, m->hstream);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(m->h_stage, m->d_lnp, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, m->hstream));
        if (grad)
            CUDA_TRY(cudaMemcpyAsync(m->h_stage + n, m->d_grad, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, m->hstream));
        CUDA_TRY(cudaStreamSynchronize(m->hstream));
        memcpy(lnp, m->h_stage, (size_t)n * sizeof(float));
        if (grad) memcpy(grad, m->h_stage + n, (size_t)n * row_bytes);
        return LINNA_OK;
    }
    StageJob job;
    job.stage_in = !is_pinned_host(u);
    job.stage_out = !is_pinned_host(lnp) || (grad && !is_pinned_host(grad));
    job.u = u, job.n = n, job.n_in = n_in, job.lnp = lnp, job.grad = grad, job.s_lnp = lnp, job.s_grad = grad;
    const float *s_in = u;
    if (job.stage_out) {
        if ((rc = ensure_pinned(&m->h_stage, &m->h_stage_cap, (size_t)n * (1 + (grad ? n_in : 0))))) return rc;
        job.s_lnp = m->h_stage, job.s_grad = m->h_stage + n;
    }
    if (job.stage_in) {
        if ((rc = ensure_pinned(&m->h_in_stage, &m->h_in_stage_cap, (size_t)n * n_in))) return rc;
        s_in = job.in_stage = m->h_in_stage;
    }
    float *const s_lnp = const_cast<float *>(job.s_lnp), *const s_grad = const_cast<float *>(job.s_grad);
    // Chunks: the kernel's time is a staircase in ROUNDS (one walker pair of 256 rows on every CTA pair: 0.17 ms at C3;
    // two interleaved pairs 0.32 ms), so chunks are cut at round boundaries and cost no extra rounds: the first chunk is
    // one round (the GPU starts after 1/6 of a 10^5-walker block has been staged), the others two rounds (at most 32
    // chunks).
    const int64_t round = (int64_t)(m->num_sms / 2) * 256;
    std::vector<int64_t> cstart{0};
    if (n > round + round / 4) {
        const int64_t big_chunk = round * std::max<int64_t>(2, (n / round + 30) / 31);
        cstart.push_back(round);
        while (cstart.back() < n) cstart.push_back(std::min(n, cstart.back() + big_chunk));
    } else
        cstart.push_back(n);
    const int nchunks = (int)cstart.size() - 1;
    job.nchunks = nchunks, job.cstart = cstart.data();
    while ((int)m->pipe_events.size() < 3 * nchunks) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        m->pipe_events.push_back(e);
    }
    job.events = m->pipe_events.data();
    // Input pieces of 512 KB, claimed front to back by the calling thread and the helper (scratch/e2e_probe.py on the
    // 16-thread GPU box: one core copies 16.5 GB/s -- 0.73 ms for 12 MB, as long as the kernel itself, so a second copier
    // is needed for the copies to hide behind the kernel; the earlier pool of 4 / 8 / 15 threads fed through a job queue
    // was SLOWER than one thread, 66 / 60 / 40 M against 71 M evals/s: wake-ups and spinning waiters cost more than the
    // copies).
    job.piece_rows = std::max<int64_t>(1, (int64_t)(((size_t)512 << 10) / row_bytes));
    job.npieces = job.stage_in ? (int)((n + job.piece_rows - 1) / job.piece_rows) : 0;
    std::vector<std::atomic<uint8_t>> piece_done(job.npieces);
    for (auto &d : piece_done) d.store(0, std::memory_order_relaxed);
    job.piece_done = piece_done.data();
    const bool use_helper = (job.stage_in || job.stage_out) && (size_t)n * row_bytes >= ((size_t)1 << 20);
    if (use_helper) {
        if (!m->helper) m->helper = new StageHelper(m->device, stage_helper_run);
        m->helper->post(&job);
    }
    rc = LINNA_OK;
    int next_h2d = 0;   // pieces [0, next_h2d) are on their way to the device
    for (int k = 0; k < nchunks && rc == LINNA_OK; ++k) {
        const int64_t r0 = cstart[k], rows = cstart[k + 1] - r0;
        cudaEvent_t landed = m->pipe_events[3 * k], done = m->pipe_events[3 * k + 1], home = m->pipe_events[3 * k + 2];
        cudaError_t ce = cudaSuccess;
        if (job.stage_in) {
            // every piece that overlaps this chunk: stage what nobody has claimed yet, wait for the helper's pieces, and
            // send runs of finished pieces to the device (pieces straddle chunk boundaries; a piece is sent once)
            const int q_end = (int)((cstart[k + 1] + job.piece_rows - 1) / job.piece_rows);
            while (next_h2d < q_end && ce == cudaSuccess) {
                if (!piece_done[next_h2d].load(std::memory_order_acquire)) {
                    const int q = job.next_piece.load(std::memory_order_relaxed) < job.npieces
                                      ? job.next_piece.fetch_add(1, std::memory_order_relaxed) : job.npieces;
                    if (q < job.npieces) stage_piece(&job, q);
                    else cpu_relax();   // the helper is finishing it
                    continue;
                }
                int q1 = next_h2d + 1;
                while (q1 < q_end && piece_done[q1].load(std::memory_order_acquire)) ++q1;
                const int64_t a0 = (int64_t)next_h2d * job.piece_rows, a1 = std::min(n, (int64_t)q1 * job.piece_rows);
                ce = cudaMemcpyAsync(m->d_in + a0 * n_in, m->h_in_stage + a0 * n_in, (size_t)(a1 - a0) * row_bytes, cudaMemcpyHostToDevice,
                                     m->cstream);
                next_h2d = q1;
            }
        } else {
            ce = cudaMemcpyAsync(m->d_in + r0 * n_in, s_in + r0 * n_in, (size_t)rows * row_bytes, cudaMemcpyHostToDevice, m->cstream);
        }
        if (ce == cudaSuccess) ce = cudaEventRecord(landed, m->cstream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(m->hstream, landed, 0);
        if (ce != cudaSuccess) { rc = fail(LINNA_ECUDA, "host->device stage: %s", cudaGetErrorString(ce)); break; }
        rc = grad ? linna_lnp_grad(m, m->d_in + r0 * n_in, rows, m->d_lnp + r0, m->d_grad + r0 * n_in, m->hstream)
                  : linna_lnp(m, m->d_in + r0 * n_in, rows, m->d_lnp + r0, m->hstream);
        if (rc) break;
        ce = cudaEventRecord(done, m->hstream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(m->dstream, done, 0);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync(s_lnp + r0, m->d_lnp + r0, (size_t)rows * sizeof(float), cudaMemcpyDeviceToHost, m->dstream);
        if (ce == cudaSuccess && grad)
            ce = cudaMemcpyAsync(s_grad + r0 * n_in, m->d_grad + r0 * n_in, (size_t)rows * row_bytes, cudaMemcpyDeviceToHost, m->dstream);
        if (ce == cudaSuccess) ce = cudaEventRecord(home, m->dstream);
        if (ce != cudaSuccess) { rc = fail(LINNA_ECUDA, "device->host stage: %s", cudaGetErrorString(ce)); break; }
        job.issued.store(k + 1, std::memory_order_release);
    }
    if (rc != LINNA_OK) job.abort.store(1);
    if (use_helper) {   // the job lives on this stack frame: the helper must be through with it
        while (!job.helper_done.load(std::memory_order_acquire)) cpu_relax();
        if (rc == LINNA_OK && job.helper_err.load()) rc = fail(LINNA_ECUDA, "device->host stage: helper thread");
    } else if (rc == LINNA_OK && job.stage_out) {
        for (int k = 0; k < nchunks; ++k) {
            const cudaError_t ce = cudaEventSynchronize(m->pipe_events[3 * k + 2]);
            if (ce != cudaSuccess) { rc = fail(LINNA_ECUDA, "device->host stage: %s", cudaGetErrorString(ce)); break; }
            copy_out_chunk(&job, k);
        }
    }
    // nothing of this call may still be in flight when it returns (or fails): the staging buffers belong to the model
    const cudaError_t e1 = cudaStreamSynchronize(m->cstream), e2 = cudaStreamSynchronize(m->hstream), e3 = cudaStreamSynchronize(m->dstream);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(LINNA_ECUDA, "host-buffer pipeline: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
    return LINNA_OK;
}

int linna_lnp_host(linna_model_t *m, const float *u, int64_t n, float *lnp)
{
    if (!m) return fail(LINNA_EINVAL, "null model");
    if (n <= 0) return n == 0 ? LINNA_OK : fail(LINNA_EINVAL, "negative n");
    if (!u || !lnp) return fail(LINNA_EINVAL, "null buffer");
    return lnp_host_pipelined(m, u, n, lnp, nullptr);
}

int linna_lnp_grad_host(linna_model_t *m, const float *u, int64_t n, float *lnp, float *grad)
{
    if (!m) return fail(LINNA_EINVAL, "null model");
    if (n <= 0) return n == 0 ? LINNA_OK : fail(LINNA_EINVAL, "negative n");
    if (!u || !lnp || !grad) return fail(LINNA_EINVAL, "null buffer");
    return lnp_host_pipelined(m, u, n, lnp, grad);
}

// ------------------------------------------------------------------------------------------ training
int linna_train_setup(linna_model_t *m, const linna_train_desc_t *d)
{
    if (!m || !d || !d->data_hat || !d->icov_hat) return fail(LINNA_EINVAL, "null argument");
    if (d->max_batch <= 0) return fail(LINNA_EINVAL, "max_batch must be > 0");
    if (m->has_extra) return fail(LINNA_EINVAL, "training the ChtoModelv2_linear variant is not supported");
    const OpHost &last = m->ops.back();
    if (last.kind != LINNA_OP_LINEAR || last.act != LINNA_ACT_NONE)
        return fail(LINNA_EINVAL, "training needs a LINEAR last layer without activation");
    const int n = m->n_out;
    m->data_hat.assign(d->data_hat, d->data_hat + n);
    m->icov_hat.assign(d->icov_hat, d->icov_hat + (size_t)n * n);
    for (int i = 0; i < n; ++i)   // x^T A x == x^T sym(A) x and the gradient is then 2 sym(A) x
        for (int j = i + 1; j < n; ++j) {
            float v = 0.5f * (m->icov_hat[(size_t)i * n + j] + m->icov_hat[(size_t)j * n + i]);
            m->icov_hat[(size_t)i * n + j] = m->icov_hat[(size_t)j * n + i] = v;
        }
    m->max_batch = d->max_batch;
    m->has_train = true;
    return rebuild(m);
}

int64_t linna_train_num_params(const linna_model_t *m) { return m ? m->n_params : -1; }

int linna_train_chisq(linna_model_t *m, const float *X, const float *Y, int64_t n, int32_t kind, float *chi2, void *stream)
{
    if (!m || !m->has_train) return fail(LINNA_ESTATE, "linna_train_setup has not been called");
    if (kind < 0 || kind > 2) return fail(LINNA_EINVAL, "bad kind");
    if (n > 0 && (!Y || !chi2)) return fail(LINNA_EINVAL, "null buffer");
    if (m->train_path == 2 && !m->tg) return fail(LINNA_EINVAL, "tensor-core training path unavailable: %s", m->tg_why.c_str());
    if (m->tg && m->train_path != 1 && n > 0) {
        if (!X) return fail(LINNA_EINVAL, "null input");
        LINNA_ON_DEVICE(m);
        cudaStream_t st = (cudaStream_t)stream;
        if (m->have_last && m->last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, m->last_done, 0));
        const int l = tg_chisq(m, m->tg, X, Y, n, kind, chi2, st);
        if (l < 0) return fail(LINNA_ECUDA, "tensor-core chi^2 launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        g_launches.fetch_add(l);
        m->last_train_kernel = 2;
        CUDA_TRY(cudaEventRecord(m->last_done, st));
        m->last_stream = st, m->have_last = true;
        return LINNA_OK;
    }
    m->last_train_kernel = 1;
    KernelArgs p;
    memset(&p, 0, sizeof p);
    p.target = Y, p.delta_kind = kind, p.loss_inv_B = 1.f;
    return run(m, PROG_LOSS, X, n, nullptr, 0, chi2, nullptr, 1, (cudaStream_t)stream, &p);
}

static AdamArgs adam_args(linna_model *m, float *params, float *am, float *av, float *grads, int64_t step, float lr,
                          float b1, float b2, float eps, float wd, int fuse)
{
    AdamArgs a;
    memset(&a, 0, sizeof a);
    a.params = params, a.m = am, a.v = av, a.grads = grads, a.blob = m->blob;
    a.map_fwd = m->map_fwd_dev, a.map_bwd = m->map_bwd_dev;
    a.lr = lr, a.beta1 = b1, a.beta2 = b2, a.eps = eps, a.wd = wd;
    a.bc1 = (float)(1.0 - std::pow((double)b1, (double)step));
    a.bc2_sqrt = (float)std::sqrt(1.0 - std::pow((double)b2, (double)step));
    a.fuse = fuse;
    return a;
}

int linna_train_step(linna_model_t *m, const float *X, const float *Y, const float *cmd, int64_t B, float *params,
                     float *adam_m, float *adam_v, float *grads, int64_t step, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int32_t fuse_adam, float *loss_rows, float *loss_mean, void *stream)
{
    if (!m || !m->has_train) return fail(LINNA_ESTATE, "linna_train_setup has not been called");
    if (B <= 0 || B > m->max_batch) return fail(LINNA_EINVAL, "batch %lld outside (0, max_batch=%d]", (long long)B, m->max_batch);
    if (!X || !Y || !cmd || !loss_rows) return fail(LINNA_EINVAL, "null buffer");
    if (fuse_adam ? (!params || !adam_m || !adam_v) : !grads) return fail(LINNA_EINVAL, "null optimiser buffer");
    if (step < 1) return fail(LINNA_EINVAL, "step counts from 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (m->train_path == 2 && !m->tg) return fail(LINNA_EINVAL, "tensor-core training path unavailable: %s", m->tg_why.c_str());
    if (m->tg && m->train_path != 1) {
        // tensor-core path: one launch per layer (forward, loss, backward-data), one for every weight gradient + AdamW
        LINNA_ON_DEVICE(m);
        if (m->have_last && m->last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, m->last_done, 0));
        AdamArgs a = adam_args(m, params, adam_m, adam_v, grads, step, lr, beta1, beta2, eps, weight_decay, fuse_adam ? 1 : 0);
        const int l = tg_train_step(m, m->tg, X, Y, cmd, B, a, loss_rows, loss_mean, st, false);
        if (l < 0) return fail(LINNA_ECUDA, "tensor-core training launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        g_launches.fetch_add(l);
        m->last_train_kernel = 2;
        CUDA_TRY(cudaEventRecord(m->last_done, st));
        m->last_stream = st, m->have_last = true;
        return LINNA_OK;
    }
    m->last_train_kernel = 1;
    KernelArgs p;
    memset(&p, 0, sizeof p);
    p.target = Y, p.cmd = cmd, p.rm_base = m->rm, p.loss_inv_B = 1.0f / (float)B, p.delta_kind = 0;
    int rc = run(m, PROG_TRAIN, X, B, nullptr, 0, loss_rows, nullptr, 1, st, &p);
    if (rc) return rc;
    AdamArgs a = adam_args(m, params, adam_m, adam_v, grads, step, lr, beta1, beta2, eps, weight_decay, fuse_adam ? 1 : 0);
    CUDA_TRY(launch_wgrad(m->wg_layers_dev, m->wg_tiles_dev, m->n_wg_tiles, m->rm, (int)B, a, st));
    g_launches.fetch_add(1);
    if (loss_mean) {
        CUDA_TRY(launch_mean(loss_rows, (int)B, loss_mean, st));
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaEventRecord(m->last_done, st));
    m->last_stream = st, m->have_last = true;
    return LINNA_OK;
}

int linna_train_set_path(linna_model_t *m, int32_t path)
{
    if (!m || path < 0 || path > 2) return fail(LINNA_EINVAL, "path must be 0 (auto), 1 (FFMA) or 2 (tensor core)");
    if (path == 2 && m->has_train && !m->tg) return fail(LINNA_EINVAL, "tensor-core training path unavailable: %s", m->tg_why.c_str());
    m->train_path = path;
    return LINNA_OK;
}

int linna_train_last_kernel(const linna_model_t *m) { return m ? m->last_train_kernel : 0; }

int linna_train_adamw(linna_model_t *m, float *params, float *adam_m, float *adam_v, const float *grads, int64_t step,
                      float lr, float beta1, float beta2, float eps, float weight_decay, void *stream)
{
    if (!m || !m->has_train) return fail(LINNA_ESTATE, "linna_train_setup has not been called");
    if (!params || !adam_m || !adam_v || !grads) return fail(LINNA_EINVAL, "null buffer");
    LINNA_ON_DEVICE(m);
    cudaStream_t st = (cudaStream_t)stream;
    if (m->have_last && m->last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, m->last_done, 0));
    AdamArgs a = adam_args(m, params, adam_m, adam_v, const_cast<float *>(grads), step, lr, beta1, beta2, eps, weight_decay, 1);
    CUDA_TRY(launch_adamw(a, (int)m->n_params, m->num_sms, st));
    g_launches.fetch_add(1);
    if (m->tg) {   // the tensor-core kernels' packed planes follow the flat vector
        CUDA_TRY(tg_repack(m->tg, params, st));
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaEventRecord(m->last_done, st));
    m->last_stream = st, m->have_last = true;
    return LINNA_OK;
}

int linna_train_adamw_peer(linna_model_t *m, float *params, float *adam_m, float *adam_v, const void *peer_grad_ptrs, int64_t grad_offset,
                           int64_t avg_offset, const void *signal_pad_ptrs, int32_t signal_slot, int32_t world, int32_t rank, int64_t step, float lr,
                           float beta1, float beta2, float eps, float weight_decay, void *stream)
{
    if (!m || !m->has_train) return fail(LINNA_ESTATE, "linna_train_setup has not been called");
    if (!params || !adam_m || !adam_v || !peer_grad_ptrs || !signal_pad_ptrs) return fail(LINNA_EINVAL, "null buffer");
    if (world < 1 || world > 16 || rank < 0 || rank >= world || grad_offset < 0 || signal_slot < 0) return fail(LINNA_EINVAL, "bad peer geometry");
    LINNA_ON_DEVICE(m);
    cudaStream_t st = (cudaStream_t)stream;
    if (m->have_last && m->last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, m->last_done, 0));
    AdamArgs a = adam_args(m, params, adam_m, adam_v, nullptr, step, lr, beta1, beta2, eps, weight_decay, 1);
    PeerReduce pr;
    if (avg_offset >= 0 && !m->peer_ticket) {
        CUDA_TRY(cudaMalloc(&m->peer_ticket, sizeof(int32_t)));
        CUDA_TRY(cudaMemset(m->peer_ticket, 0, sizeof(int32_t)));
    }
    pr.grads = reinterpret_cast<const float *const *>(peer_grad_ptrs), pr.offset = grad_offset;
    pr.avg_offset = avg_offset, pr.ticket = m->peer_ticket;
    pr.pads = reinterpret_cast<uint32_t *const *>(signal_pad_ptrs), pr.slot = signal_slot;
    pr.world = world, pr.rank = rank, pr.token = ++m->peer_token;
    CUDA_TRY(launch_adamw_peer(a, (int)m->n_params, m->num_sms, pr, st));
    g_launches.fetch_add(1);
    if (m->tg) {   // the tensor-core kernels' packed planes follow the flat vector
        CUDA_TRY(tg_repack(m->tg, params, st));
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaEventRecord(m->last_done, st));
    m->last_stream = st, m->have_last = true;
    return LINNA_OK;
}

int linna_train_load_params(linna_model_t *m, const float *params, void *stream)
{
    if (!m || !m->has_train) return fail(LINNA_ESTATE, "linna_train_setup has not been called");
    if (!params) return fail(LINNA_EINVAL, "null buffer");
    LINNA_ON_DEVICE(m);
    cudaStream_t st = (cudaStream_t)stream;
    if (m->have_last && m->last_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, m->last_done, 0));
    CUDA_TRY(launch_scatter_params(params, m->blob, m->map_fwd_dev, m->map_bwd_dev, (int)m->n_params, st));
    g_launches.fetch_add(1);
    if (m->tg) {
        CUDA_TRY(tg_repack(m->tg, params, st));
        g_launches.fetch_add(1);
    }
    CUDA_TRY(cudaEventRecord(m->last_done, st));
    m->last_stream = st, m->have_last = true;
    return LINNA_OK;
}

int linna_train_commit(linna_model_t *m, const float *ph)
{
    if (!m || !ph) return fail(LINNA_EINVAL, "null argument");
    size_t o = 0;
    auto take = [&](std::vector<float> &v) { std::copy(ph + o, ph + o + v.size(), v.begin()); o += v.size(); };
    for (OpHost &op : m->ops) {
        take(op.w), take(op.b);
        if (op.kind == LINNA_OP_RES) {
            take(op.w2), take(op.b2);
            if (op.has_ws) take(op.ws);
        }
    }
    return rebuild(m);
}

}  // extern "C"
