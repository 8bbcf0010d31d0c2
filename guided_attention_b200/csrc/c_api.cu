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
        cudaStream_t);
int bwd(const void*, const void*, const void*, const float*, const void*, const float*, int64_t, int, void*, float*,
        float*, int, int, int, int, int, float, int, cudaStream_t);
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

extern "C" int ga_cross_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, float* acc,
                                 int batch, int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype,
                                 int impl, ga_stream_t stream) {
  int rc = check_attn_args(q, k, batch, heads, n_query, n_ctx, head_dim, dtype);
  if (rc != GA_OK) return rc;
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
  return simt::fwd(q, k, v, o, lse, acc, nullptr, batch, heads, n_query, n_ctx, head_dim, scale, dtype, st);
}

extern "C" int ga_cross_attn_bwd(const void* q, const void* k, const void* v, const float* lse, const void* d_o,
                                 const float* d_acc, int64_t d_acc_batch_stride, int d_acc_row_stride, void* d_q,
                                 float* d_k, float* d_v, int batch, int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype,
                                 int impl, ga_stream_t stream) {
  int rc = check_attn_args(q, k, batch, heads, n_query, n_ctx, head_dim, dtype);
  if (rc != GA_OK) return rc;
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
                   n_ctx, head_dim, scale, dtype, st);
}

extern "C" int ga_attn_probs(const void* q, const void* k, void* probs, int batch, int heads, int n_query, int n_ctx,
                             int head_dim, float scale, int dtype, ga_stream_t stream) {
  int rc = check_attn_args(q, k, batch, heads, n_query, n_ctx, head_dim, dtype);
  if (rc != GA_OK) return rc;
  GA_CHECK_ARG(probs != nullptr, "probs is NULL");
  return simt::fwd(q, k, k, nullptr, nullptr, nullptr, probs, batch, heads, n_query, n_ctx, head_dim, scale, dtype,
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
