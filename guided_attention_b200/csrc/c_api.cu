// Library-level entry points of the C ABI and the variant dispatch for the attention kernels (include/guided_attn.h).
#include "ga_common.cuh"

namespace ga {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

namespace simt {
int fwd(const void*, const void*, const void*, void*, float*, float*, void*, int, int, int, int, int, float, int,
        const ga_score_bias_t*, cudaStream_t);
int bwd(const void*, const void*, const void*, const float*, const void*, const float*, int64_t, int, void*, float*,
        float*, int, int, int, int, int, float, int, const ga_score_bias_t*, float*, cudaStream_t);
int smax(const void*, const void*, unsigned long long*, int, int, int, int, int, float, int, const ga_score_bias_t*,
         cudaStream_t);
}  // namespace simt
namespace tc {
bool supports_fwd(int dtype, int n_ctx, int head_dim, int heads, bool with_acc);
bool supports_bwd(int dtype, int n_ctx, int head_dim, int heads, bool with_dkv);
int fwd(const void*, const void*, const void*, void*, float*, float*, int, int, int, int, int, float, int, int,
        cudaStream_t);
int bwd(const void*, const void*, const void*, const float*, const void*, const float*, int64_t, int, void*, int, int,
        int, int, int, float, int, int, cudaStream_t);
}  // namespace tc

namespace sa {
bool supports(int dtype, int head_dim);
int fwd(const void*, const void*, const void*, void*, float*, int, int, int, int, float, int, cudaStream_t);
int bwd(const void*, const void*, const void*, const void*, const float*, const void*, void*, void*, void*, float*, int,
        int, int, int, float, int, cudaStream_t);
}  // namespace sa

static int check_attn_args(const void* q, const void* k, int batch, int heads, int n_query, int n_ctx, int head_dim,
                           int dtype) {
  GA_CHECK_ARG(q != nullptr && k != nullptr, "NULL operand");
  GA_CHECK_ARG(batch >= 1 && heads >= 1 && n_query >= 1, "bad batch %d / heads %d / n_query %d", batch, heads, n_query);
  GA_CHECK_ARG(n_ctx >= 1 && n_ctx <= GA_MAX_CTX, "n_ctx %d out of range [1, %d]", n_ctx, GA_MAX_CTX);
  GA_CHECK_ARG(head_dim >= 8 && head_dim <= 256 && head_dim % 8 == 0, "head_dim %d must be a multiple of 8 in [8, 256]",
               head_dim);
  GA_CHECK_ARG(dtype == GA_F32 || dtype == GA_F16 || dtype == GA_BF16, "unknown dtype %d", dtype);
  GA_CHECK_ALIGN(q, 16, "q");
  GA_CHECK_ALIGN(k, 16, "k");
  return GA_OK;
}

}  // namespace ga

using namespace ga;

extern "C" int ga_version(void) { return GA_ABI_VERSION; }
extern "C" const char* ga_last_error(void) { return g_err; }

extern "C" int ga_device_supported(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(GA_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  return prop.major == 10 ? 1 : 0;
}

static int check_bias(const ga_score_bias_t* b, int batch, int heads, int n_query, int n_ctx, int impl) {
  if (b == nullptr) return GA_OK;
  GA_CHECK_ARG(impl == GA_IMPL_AUTO || impl == GA_IMPL_SIMT, "the score bias is implemented by the SIMT variant only");
  GA_CHECK_ARG(b->pww_count >= 0 && b->pww_count <= GA_MAX_TOKENS, "pww_count %d out of range", b->pww_count);
  if (b->pww_count > 0) {
    GA_CHECK_ARG(b->pww_masks && b->pww_coef && b->pww_smax, "paint-with-words needs pww_masks, pww_coef and pww_smax");
    GA_CHECK_ARG((int64_t)batch * heads * n_query * n_ctx < (int64_t)0xffffffffll, "launch too large for the packed max");
    GA_CHECK_ALIGN(b->pww_smax, 8, "pww_smax");
    for (int i = 0; i < b->pww_count; ++i)
      GA_CHECK_ARG(b->pww_column[i] >= 0 && b->pww_column[i] < n_ctx, "pww_column[%d] = %d", i, b->pww_column[i]);
  }
  return GA_OK;
}

extern "C" int ga_cross_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, float* acc,
                                 int batch, int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype,
                                 int impl, ga_stream_t stream) {
  return ga_cross_attn_fwd_ex(q, k, v, o, lse, acc, nullptr, batch, heads, n_query, n_ctx, head_dim, scale, dtype, impl,
                              stream);
}

extern "C" int ga_cross_attn_fwd_ex(const void* q, const void* k, const void* v, void* o, float* lse, float* acc,
                                    const ga_score_bias_t* bias, int batch, int heads, int n_query, int n_ctx,
                                    int head_dim, float scale, int dtype, int impl, ga_stream_t stream) {
  int rc = check_attn_args(q, k, batch, heads, n_query, n_ctx, head_dim, dtype);
  if (rc != GA_OK) return rc;
  if ((rc = check_bias(bias, batch, heads, n_query, n_ctx, impl)) != GA_OK) return rc;
  if (bias != nullptr) impl = GA_IMPL_SIMT;
  GA_CHECK_ARG(v != nullptr && o != nullptr && lse != nullptr, "NULL operand");
  GA_CHECK_ALIGN(v, 16, "v");
  GA_CHECK_ALIGN(o, 16, "o");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool tc_ok = tc::supports_fwd(dtype, n_ctx, head_dim, heads, acc != nullptr);
  const bool want_tc = impl == GA_IMPL_TCGEN05 || impl == GA_IMPL_TCGEN05_SINGLE || impl == GA_IMPL_TCGEN05_PIPE;
  if (want_tc && !tc_ok)
    return fail(GA_ERR_UNSUPPORTED, "tcgen05 cross-attention does not support dtype %d / n_ctx %d / head_dim %d", dtype,
                n_ctx, head_dim);
  if (want_tc || (impl == GA_IMPL_AUTO && tc_ok)) {
    const int force = impl == GA_IMPL_TCGEN05_SINGLE ? 0 : (impl == GA_IMPL_TCGEN05_PIPE ? 1 : -1);
    return tc::fwd(q, k, v, o, lse, acc, batch, heads, n_query, n_ctx, head_dim, scale, dtype, force, st);
  }
  return simt::fwd(q, k, v, o, lse, acc, nullptr, batch, heads, n_query, n_ctx, head_dim, scale, dtype, bias, st);
}

extern "C" int ga_cross_attn_smax(const void* q, const void* k, unsigned long long* smax, const ga_score_bias_t* bias,
                                  int batch, int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype,
                                  ga_stream_t stream) {
  int rc = check_attn_args(q, k, batch, heads, n_query, n_ctx, head_dim, dtype);
  if (rc != GA_OK) return rc;
  GA_CHECK_ARG(smax != nullptr, "smax is NULL");
  GA_CHECK_ALIGN(smax, 8, "smax");
  GA_CHECK_ARG((int64_t)batch * heads * n_query * n_ctx < (int64_t)0xffffffffll, "launch too large for the packed max");
  return simt::smax(q, k, smax, batch, heads, n_query, n_ctx, head_dim, scale, dtype, bias,
                    static_cast<cudaStream_t>(stream));
}

extern "C" int ga_cross_attn_bwd(const void* q, const void* k, const void* v, const float* lse, const void* d_o,
                                 const float* d_acc, int64_t d_acc_batch_stride, int d_acc_row_stride, void* d_q,
                                 float* d_k, float* d_v, int batch, int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype,
                                 int impl, ga_stream_t stream) {
  return ga_cross_attn_bwd_ex(q, k, v, lse, d_o, d_acc, d_acc_batch_stride, d_acc_row_stride, d_q, d_k, d_v, nullptr,
                              nullptr, batch, heads, n_query, n_ctx, head_dim, scale, dtype, impl, stream);
}

extern "C" int ga_cross_attn_bwd_ex(const void* q, const void* k, const void* v, const float* lse, const void* d_o,
                                    const float* d_acc, int64_t d_acc_batch_stride, int d_acc_row_stride, void* d_q,
                                    float* d_k, float* d_v, const ga_score_bias_t* bias, float* pww_partials, int batch,
                                    int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype, int impl,
                                    ga_stream_t stream) {
  int rc = check_attn_args(q, k, batch, heads, n_query, n_ctx, head_dim, dtype);
  if (rc != GA_OK) return rc;
  if ((rc = check_bias(bias, batch, heads, n_query, n_ctx, impl)) != GA_OK) return rc;
  if (bias != nullptr) {
    impl = GA_IMPL_SIMT;
    GA_CHECK_ARG(bias->pww_count == 0 || pww_partials != nullptr, "paint-with-words backward needs pww_partials");
  }
  GA_CHECK_ARG(v != nullptr && lse != nullptr && d_o != nullptr && d_q != nullptr, "NULL operand");
  GA_CHECK_ALIGN(v, 16, "v");
  GA_CHECK_ALIGN(d_o, 16, "d_o");
  GA_CHECK_ALIGN(d_q, 16, "d_q");
  GA_CHECK_ARG(d_acc == nullptr || d_acc_row_stride >= n_ctx, "d_acc_row_stride %d < n_ctx %d", d_acc_row_stride, n_ctx);
  if (d_acc != nullptr && (d_acc_row_stride & 3) == 0) GA_CHECK_ALIGN(d_acc, 16, "d_acc (16-byte rows)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool want_dkv = d_k != nullptr || d_v != nullptr;
  const bool tc_ok = tc::supports_bwd(dtype, n_ctx, head_dim, heads, want_dkv);
  const bool want_tc = impl == GA_IMPL_TCGEN05 || impl == GA_IMPL_TCGEN05_SINGLE || impl == GA_IMPL_TCGEN05_PIPE;
  if (want_tc && !tc_ok)
    return fail(GA_ERR_UNSUPPORTED, "tcgen05 cross-attention backward does not support this configuration");
  if (want_tc || (impl == GA_IMPL_AUTO && tc_ok)) {
    const int force = impl == GA_IMPL_TCGEN05_SINGLE ? 0 : (impl == GA_IMPL_TCGEN05_PIPE ? 1 : -1);
    return tc::bwd(q, k, v, lse, d_o, d_acc, d_acc_batch_stride, d_acc_row_stride, d_q, batch, heads, n_query, n_ctx,
                   head_dim, scale, dtype, force, st);
  }
  return simt::bwd(q, k, v, lse, d_o, d_acc, d_acc_batch_stride, d_acc_row_stride, d_q, d_k, d_v, batch, heads, n_query,
                   n_ctx, head_dim, scale, dtype, bias, pww_partials, st);
}

extern "C" int ga_attn_probs(const void* q, const void* k, void* probs, int batch, int heads, int n_query, int n_ctx,
                             int head_dim, float scale, int dtype, ga_stream_t stream) {
  return ga_attn_probs_ex(q, k, probs, nullptr, batch, heads, n_query, n_ctx, head_dim, scale, dtype, stream);
}

extern "C" int ga_attn_probs_ex(const void* q, const void* k, void* probs, const ga_score_bias_t* bias, int batch,
                                int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype,
                                ga_stream_t stream) {
  int rc = check_attn_args(q, k, batch, heads, n_query, n_ctx, head_dim, dtype);
  if (rc != GA_OK) return rc;
  if ((rc = check_bias(bias, batch, heads, n_query, n_ctx, GA_IMPL_SIMT)) != GA_OK) return rc;
  GA_CHECK_ARG(probs != nullptr, "probs is NULL");
  return simt::fwd(q, k, k, nullptr, nullptr, nullptr, probs, batch, heads, n_query, n_ctx, head_dim, scale, dtype, bias,
                   static_cast<cudaStream_t>(stream));
}

extern "C" int ga_self_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int batch, int heads,
                                int n_tokens, int head_dim, float scale, int dtype, ga_stream_t stream) {
  GA_CHECK_ARG(q && k && v && o && lse, "NULL operand");
  GA_CHECK_ARG(batch >= 1 && heads >= 1 && n_tokens >= 1, "bad batch %d / heads %d / n_tokens %d", batch, heads, n_tokens);
  if (!sa::supports(dtype, head_dim))
    return fail(GA_ERR_UNSUPPORTED, "self-attention kernel: dtype %d / head_dim %d not supported", dtype, head_dim);
  GA_CHECK_ALIGN(q, 16, "q"); GA_CHECK_ALIGN(k, 16, "k"); GA_CHECK_ALIGN(v, 16, "v"); GA_CHECK_ALIGN(o, 16, "o");
  return sa::fwd(q, k, v, o, lse, batch, heads, n_tokens, head_dim, scale, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int ga_self_attn_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse,
                                const void* d_o, void* d_q, void* d_k, void* d_v, float* dvec, int batch, int heads,
                                int n_tokens, int head_dim, float scale, int dtype, ga_stream_t stream) {
  GA_CHECK_ARG(q && k && v && o && lse && d_o && d_q && d_k && d_v && dvec, "NULL operand");
  GA_CHECK_ARG(batch >= 1 && heads >= 1 && n_tokens >= 1, "bad batch %d / heads %d / n_tokens %d", batch, heads, n_tokens);
  if (!sa::supports(dtype, head_dim))
    return fail(GA_ERR_UNSUPPORTED, "self-attention kernel: dtype %d / head_dim %d not supported", dtype, head_dim);
  GA_CHECK_ALIGN(q, 16, "q"); GA_CHECK_ALIGN(k, 16, "k"); GA_CHECK_ALIGN(v, 16, "v"); GA_CHECK_ALIGN(o, 16, "o");
  GA_CHECK_ALIGN(d_o, 16, "d_o"); GA_CHECK_ALIGN(d_q, 16, "d_q"); GA_CHECK_ALIGN(d_k, 16, "d_k");
  GA_CHECK_ALIGN(d_v, 16, "d_v");
  return sa::bwd(q, k, v, o, lse, d_o, d_q, d_k, d_v, dvec, batch, heads, n_tokens, head_dim, scale, dtype,
                 static_cast<cudaStream_t>(stream));
}
