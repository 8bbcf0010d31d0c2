// tcgen05 / TMEM / TMA / mbarrier PTX wrappers and UMMA descriptor helpers shared by the tensor-core kernels
// (cross_attn_tc.cu, self_attn_tc.cu).  sm_100a only.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "ga_common.cuh"

namespace ga {
namespace tc {

constexpr uint32_t kSpinLimit = 1u << 24;
constexpr int kTpad = 80;       // cross-attention: keys padded to a multiple of 16 (UMMA N granularity at M = 128)

// ----------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure reported to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < kSpinLimit; ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors (cute/arch/mma_sm100_desc.hpp bit layout) -------------------------------------------------
// shared-memory matrix descriptor, 128-byte swizzle: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1
// [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32 [4,6)=1 | A fmt [7,10) | B fmt [10,13) | A major bit15 | B major bit16 | N>>3 [17,23)
// | M>>4 [24,29)      (fmt: 0 = f16, 1 = bf16; major: 0 = K, 1 = MN)
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int b_mn_major, int n, int m) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack16(float a, float b, bool bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Row softmax over up to kTpad keys held by one thread, un-normalised: on return s[j] = exp(scale*(s[j] - max)) for
// j < T and 0 beyond, packed[] holds the same values as 16-bit pairs (the A operand of the second GEMM), and the row sum
// is returned through `sum` (the normalisation 1/sum is applied to the fp32 O accumulator and to the map accumulator,
// not to the 16-bit operand).  Only the last 16 columns are predicated on T: the callers guarantee T > kTpad - 16.
__device__ __forceinline__ void row_softmax(float* s, uint32_t* packed, int T, float sc, bool bf16, float& m, float& sum) {
  m = -INFINITY;
#pragma unroll
  for (int j = 0; j < kTpad - 16; ++j) m = fmaxf(m, s[j]);
#pragma unroll
  for (int j = kTpad - 16; j < kTpad; ++j)
    if (j < T) m = fmaxf(m, s[j]);
  const float mo = m * sc;
  sum = 0.f;
#pragma unroll
  for (int j = 0; j < kTpad - 16; ++j) {
    s[j] = ex2_approx(fmaf(s[j], sc, -mo));
    sum += s[j];
  }
#pragma unroll
  for (int j = kTpad - 16; j < kTpad; ++j) {
    s[j] = (j < T) ? ex2_approx(fmaf(s[j], sc, -mo)) : 0.f;
    sum += s[j];
  }
#pragma unroll
  for (int j = 0; j < kTpad; j += 2) packed[j >> 1] = pack16(s[j], s[j + 1], bf16);
}


// ---- warp-uniform role helpers ----------------------------------------------------------------------------------
// The TMA-producer and MMA-issuer roles run as WHOLE warps in warp-uniform control flow and pick one lane with
// elect.sync for the asynchronous instructions.  A role entered through a divergent `lane == 0` branch makes the
// compiler wrap every UTCHMMA / UTMALDG in an ELECT / BRA.U.ANY loop and keep descriptors in vector registers
// (~13 SASS instructions per MMA): the single issuing thread then becomes the critical path of the whole CTA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// start-address field of a shared-memory descriptor advanced by `bytes` (stays inside the 14-bit field: < 256 KB)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// D[tmem] = A . B^T with A and B K-major in 64-column blocks of 128-byte swizzled rows; one k-step = 16 columns =
// 32 bytes inside the row, four k-steps per block.
__device__ __forceinline__ void issue_kmajor_gemm(uint32_t d_tmem, uint64_t a_desc, uint32_t a_block_bytes, uint64_t b_desc,
                                                  uint32_t b_block_bytes, int ksteps, uint32_t idesc) {
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint32_t blk = (uint32_t)ks >> 2, in = ((uint32_t)ks & 3u) * 32u;
    mma_ss(d_tmem, desc_advance(a_desc, blk * a_block_bytes + in), desc_advance(b_desc, blk * b_block_bytes + in), idesc,
           ks > 0 ? 1u : 0u);
  }
}
// D[tmem] (+)= A[tmem, packed 16-bit: 8 columns per k-step] . B with B MN-major (k-step = 16 rows of 128 bytes)
__device__ __forceinline__ void issue_tmem_gemm(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, int ksteps, uint32_t idesc,
                                                bool accumulate) {
  for (int ks = 0; ks < ksteps; ++ks)
    mma_ts(d_tmem, a_tmem + (uint32_t)ks * 8u, desc_advance(b_desc, (uint32_t)ks * 2048u), idesc,
           (accumulate || ks > 0) ? 1u : 0u);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int kRegs> __device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs> __device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}


// ---- host side --------------------------------------------------------------------------------------------------
constexpr int kBlockCols = 64;  // 16-bit elements per 128-byte swizzle row
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// (d, H, rows, B) view of a (B, rows, H*d) tensor; box = 64 channels x 1 head x box_rows rows; 128-byte swizzle;
// out-of-range channels / rows are filled with zeros.
static inline int make_map(CUtensorMap* map, const void* ptr, int dtype, int B, int rows, int H, int d, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(GA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)H, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)d * 2, (cuuint64_t)H * d * 2, (cuuint64_t)rows * H * d * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockCols, 1u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(map, dtype == GA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT) {
    // first call on a thread that has no current context yet (e.g. a fresh autograd worker): let the runtime bind the
    // primary context of the current device, then retry
    cudaFree(nullptr);
    r = fn(map, dtype == GA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
           const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return fail(GA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return GA_OK;
}

// The opt-in dynamic shared-memory limit is raised once per (device, kernel) to the largest size the kernel can ask
// for, so steady-state launches (and CUDA-graph captures) make no attribute call at all.
static inline cudaError_t ensure_smem(const void* kernel, int slot, size_t) {
  static bool done[64][8] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (done[dev][slot]) return cudaSuccess;
  int optin = 0;
  cudaFuncAttributes fa;
  if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
  if ((e = cudaFuncGetAttributes(&fa, kernel)) != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
  if (e == cudaSuccess) done[dev][slot] = true;
  return e;
}

static inline int sm_count() {
  static int n[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev] > 0 ? n[dev] : 148;
}


}  // namespace tc
}  // namespace ga
