// K1 / K2, exact-fp32 SIMT variant: cross-attention against a short text context (T <= 128 keys) with the
// attention-map accumulation fused in.  This is the variant for fp32 operands (the tensor cores would round to tf32,
// outside the 1e-3 parity budget) and the on-device cross-check for the tcgen05 variant (cross_attn_tc.cu).
//
// Replaces reference utils/ptp_utils.py:77-85, 97-146 (QK^T -> softmax -> store -> PV) and its autograd.
// Layout: q/o (B, N, H*d), k/v (B, T, H*d): head h of token n starts at ((b*N + n)*H + h)*d, d contiguous.
//
// One CTA = 64 query rows x `heads_per_cta` heads.  K, V of the head and the Q tile are staged in shared memory as
// 32-bit words (1 float or 2 halves) with an odd row stride, so that lanes that own different keys read different
// banks.  A warp owns 8 rows, processed 4 at a time: lanes are keys for S = QK^T and the softmax (row reductions are
// warp shuffles), then lanes are channels for O = PV.  The head-sum of P for the AttentionStore accumulator stays in
// registers across the head loop, so `acc` is written once, without atomics, in a fixed order (deterministic).
#include "ga_common.cuh"

namespace ga {
namespace simt {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsPerCta = 64;
constexpr int kRowsPerWarp = kRowsPerCta / kWarps;  // 8
constexpr int kR = 4;                               // rows processed together by a warp
constexpr int kGroups = kRowsPerWarp / kR;          // 2
constexpr int kKPL = GA_MAX_CTX / 32;               // keys per lane (4)
constexpr int kPStride = GA_MAX_CTX;                // per-row stride of the probability scratch
constexpr int kMaxHeadDim = 256;

template <typename T> struct Cfg {
  static constexpr int E = Word<T>::E;
  static constexpr int kWPL = kMaxHeadDim / E / 32;  // channel words per lane: 8 (fp32) or 4 (16-bit)
};

__device__ __forceinline__ int odd_stride(int words) { return words | 1; }

// Cooperative copy of `rows` rows of `words` 32-bit words (row pitch `pitch_words` in global memory) into shared
// memory with stride `stride`; rows >= valid_rows are zero-filled.  16-byte global loads.
__device__ __forceinline__ void stage_rows(uint32_t* dst, const uint32_t* src, int rows, int valid_rows, int words,
                                           int64_t pitch_words, int stride) {
  const int vec_per_row = words >> 2;  // words % 4 == 0 is checked on the host (d % 8 == 0)
  for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, c = (i - r * vec_per_row) << 2;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (r < valid_rows) val = __ldg(reinterpret_cast<const uint4*>(src + r * pitch_words + c));
    uint32_t* p = dst + r * stride + c;
    p[0] = val.x; p[1] = val.y; p[2] = val.z; p[3] = val.w;
  }
}

template <typename T>
__device__ __forceinline__ void dot_step(float& s, float2 a, float2 b) {
  s = fmaf(a.x, b.x, s);
  if (Word<T>::E == 2) s = fmaf(a.y, b.y, s);
}

// ---- optional score bias (include/guided_attn.h: ga_score_bias_t) --------------------------------------------------
// attention_mask (reference utils/ptp_utils.py:135-136) and the paint-with-words bias mask * 0.4 * max(S) * ln(1+sigma)
// (:113-138).  The global max arrives packed in one 64-bit word (ordered-float bits << 32 | ~flat index) written by
// `cross_attn_smax_kernel`, so that the backward can also find WHERE the max was (the gradient through it).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct BiasLane {
  uint32_t colmask[kKPL];   // bit i set: paint-with-words entry i biases the key column this lane owns in slot kk
  float pww;                // coef * max(S), 0 when paint-with-words is off
};

__device__ __forceinline__ BiasLane bias_lane_setup(const ga_score_bias_t& sb, int lane) {
  BiasLane bl;
#pragma unroll
  for (int kk = 0; kk < kKPL; ++kk) {
    bl.colmask[kk] = 0u;
    for (int i = 0; i < sb.pww_count; ++i)
      if (sb.pww_column[i] == lane + 32 * kk) bl.colmask[kk] |= 1u << i;
  }
  bl.pww = 0.f;
  if (sb.pww_count > 0) bl.pww = __ldg(sb.pww_coef) * ordered_to_float((uint32_t)(*sb.pww_smax >> 32));
  return bl;
}

// bit i: query row `row` lies inside box i at this layer's resolution (whole warp calls this)
__device__ __forceinline__ uint32_t pww_row_bits(const ga_score_bias_t& sb, int row, int N, int lane) {
  if (sb.pww_count == 0) return 0u;
  const bool in = lane < sb.pww_count && row < N && sb.pww_masks[(int64_t)lane * N + row] != 0;
  return __ballot_sync(0xffffffffu, in);
}

__device__ __forceinline__ float score_bias(const ga_score_bias_t& sb, const BiasLane& bl, uint32_t bits, int bh,
                                            int row, int N, int j, int kk, int Tctx) {
  float add = 0.f;
  if (sb.mask != nullptr && row < N && j < Tctx)
    add = __ldg(sb.mask + (int64_t)bh * sb.mask_stride_bh + (int64_t)row * sb.mask_stride_n + j);
  return add + bl.pww * (float)__popc(bits & bl.colmask[kk]);
}

// ------------------------------------------------------------------------------------------------------ forward
template <typename T, bool kProbsOnly, bool kBias>
__global__ void __launch_bounds__(kThreads)
cross_attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, T* __restrict__ o,
                      float* __restrict__ lse, float* __restrict__ acc, T* __restrict__ probs, int H, int N, int Tctx,
                      int d, float scale, int heads_per_cta, const ga_score_bias_t sb) {
  constexpr int E = Cfg<T>::E;
  constexpr int kWPL = Cfg<T>::kWPL;
  extern __shared__ uint32_t smem[];
  const int words = d / E;
  const int stride = odd_stride(words);
  uint32_t* sK = smem;
  uint32_t* sV = sK + Tctx * stride;
  uint32_t* sQ = sV + Tctx * stride;
  float* sP = reinterpret_cast<float*>(sQ + kRowsPerCta * stride);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kRowsPerCta;
  const int bh0 = blockIdx.y * heads_per_cta;
  const int64_t pitch = (int64_t)H * words;  // words between consecutive tokens

  float pacc[kGroups][kR][kKPL];
#pragma unroll
  for (int g = 0; g < kGroups; ++g)
#pragma unroll
    for (int r = 0; r < kR; ++r)
#pragma unroll
      for (int kk = 0; kk < kKPL; ++kk) pacc[g][r][kk] = 0.f;

  int jc[kKPL];
#pragma unroll
  for (int kk = 0; kk < kKPL; ++kk) jc[kk] = min(lane + 32 * kk, Tctx - 1);
  BiasLane bl;
  if (kBias) bl = bias_lane_setup(sb, lane);

  for (int hh = 0; hh < heads_per_cta; ++hh) {
    const int bh = bh0 + hh, b = bh / H, h = bh - b * H;
    const uint32_t* gq = reinterpret_cast<const uint32_t*>(q) + ((int64_t)b * N + row0) * pitch + (int64_t)h * words;
    const uint32_t* gk = reinterpret_cast<const uint32_t*>(k) + (int64_t)b * Tctx * pitch + (int64_t)h * words;
    const uint32_t* gv = reinterpret_cast<const uint32_t*>(v) + (int64_t)b * Tctx * pitch + (int64_t)h * words;
    __syncthreads();
    stage_rows(sK, gk, Tctx, Tctx, words, pitch, stride);
    if (!kProbsOnly) stage_rows(sV, gv, Tctx, Tctx, words, pitch, stride);
    stage_rows(sQ, gq, kRowsPerCta, N - row0, words, pitch, stride);
    __syncthreads();

#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
      const int rl = warp * kRowsPerWarp + g * kR;  // first local row of the group
      float s[kR][kKPL];
#pragma unroll
      for (int r = 0; r < kR; ++r)
#pragma unroll
        for (int kk = 0; kk < kKPL; ++kk) s[r][kk] = 0.f;
      for (int kw = 0; kw < words; ++kw) {
        float2 qv[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) qv[r] = Word<T>::unpack(sQ[(rl + r) * stride + kw]);
#pragma unroll
        for (int kk = 0; kk < kKPL; ++kk) {
          const float2 kv = Word<T>::unpack(sK[jc[kk] * stride + kw]);
#pragma unroll
          for (int r = 0; r < kR; ++r) dot_step<T>(s[r][kk], qv[r], kv);
        }
      }
      float* pw = sP + (warp * kR) * kPStride;
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const int row = row0 + rl + r;
        float m = -INFINITY;
        uint32_t bits = 0u;
        if (kBias) bits = pww_row_bits(sb, row, N, lane);
#pragma unroll
        for (int kk = 0; kk < kKPL; ++kk) {
          s[r][kk] *= scale;
          if (kBias) s[r][kk] += score_bias(sb, bl, bits, bh, row, N, lane + 32 * kk, kk, Tctx);
          if (lane + 32 * kk < Tctx) m = fmaxf(m, s[r][kk]);
        }
        m = warp_max(m);
        float e[kKPL], sum = 0.f;
#pragma unroll
        for (int kk = 0; kk < kKPL; ++kk) {
          e[kk] = (lane + 32 * kk < Tctx) ? expf(s[r][kk] - m) : 0.f;
          sum += e[kk];
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
#pragma unroll
        for (int kk = 0; kk < kKPL; ++kk) {
          const float p = e[kk] * inv;
          const int j = lane + 32 * kk;
          if (j < Tctx) {
            if (kProbsOnly) {
              if (row < N) probs[((int64_t)bh * N + row) * Tctx + j] = static_cast<T>(p);
            } else {
              pw[r * kPStride + j] = p;
            }
          }
          pacc[g][r][kk] += p;
        }
        if (!kProbsOnly && lane == 0 && row < N) lse[(int64_t)bh * N + row] = m + logf(sum);
      }
      if (kProbsOnly) continue;
      __syncwarp();

      float oacc[kR][kWPL][E];
#pragma unroll
      for (int r = 0; r < kR; ++r)
#pragma unroll
        for (int m = 0; m < kWPL; ++m)
#pragma unroll
          for (int e2 = 0; e2 < E; ++e2) oacc[r][m][e2] = 0.f;
      for (int j = 0; j < Tctx; ++j) {
        float pj[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) pj[r] = pw[r * kPStride + j];
#pragma unroll
        for (int m = 0; m < kWPL; ++m) {
          const int cw = lane + 32 * m;
          if (cw < words) {
            const float2 vv = Word<T>::unpack(sV[j * stride + cw]);
#pragma unroll
            for (int r = 0; r < kR; ++r) {
              oacc[r][m][0] = fmaf(pj[r], vv.x, oacc[r][m][0]);
              if (E == 2) oacc[r][m][E - 1] = fmaf(pj[r], vv.y, oacc[r][m][E - 1]);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const int row = row0 + rl + r;
        if (row >= N) continue;
        uint32_t* go = reinterpret_cast<uint32_t*>(o) + ((int64_t)b * N + row) * pitch + (int64_t)h * words;
#pragma unroll
        for (int m = 0; m < kWPL; ++m) {
          const int cw = lane + 32 * m;
          if (cw < words) go[cw] = Word<T>::pack(oacc[r][m][0], oacc[r][m][E - 1]);
        }
      }
      __syncwarp();
    }
  }

  if (!kProbsOnly && acc != nullptr) {
    const int b = bh0 / H;
#pragma unroll
    for (int g = 0; g < kGroups; ++g)
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const int row = row0 + warp * kRowsPerWarp + g * kR + r;
        if (row >= N) continue;
#pragma unroll
        for (int kk = 0; kk < kKPL; ++kk) {
          const int j = lane + 32 * kk;
          if (j < Tctx) acc[((int64_t)b * N + row) * Tctx + j] = pacc[g][r][kk];
        }
      }
  }
}

// ------------------------------------------------------------------------------- global score max (paint-with-words)
// max over (b, h, n, t) of scale * q.k (+ attention_mask), packed with its flat index ((b*H + h)*N + n)*T + t; ties go
// to the lowest index.  One 64-bit atomicMax per warp and row group; the result does not depend on the order.
template <typename T>
__global__ void __launch_bounds__(kThreads)
cross_attn_smax_kernel(const T* __restrict__ q, const T* __restrict__ k, unsigned long long* __restrict__ smax, int H,
                       int N, int Tctx, int d, float scale, const ga_score_bias_t sb) {
  constexpr int E = Cfg<T>::E;
  extern __shared__ uint32_t smem[];
  const int words = d / E;
  const int stride = odd_stride(words);
  uint32_t* sK = smem;
  uint32_t* sQ = sK + Tctx * stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kRowsPerCta;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int64_t pitch = (int64_t)H * words;
  stage_rows(sK, reinterpret_cast<const uint32_t*>(k) + (int64_t)b * Tctx * pitch + (int64_t)h * words, Tctx, Tctx, words,
             pitch, stride);
  stage_rows(sQ, reinterpret_cast<const uint32_t*>(q) + ((int64_t)b * N + row0) * pitch + (int64_t)h * words,
             kRowsPerCta, N - row0, words, pitch, stride);
  __syncthreads();
  int jc[kKPL];
#pragma unroll
  for (int kk = 0; kk < kKPL; ++kk) jc[kk] = min(lane + 32 * kk, Tctx - 1);
  unsigned long long best = 0ull;
#pragma unroll 1
  for (int g = 0; g < kGroups; ++g) {
    const int rl = warp * kRowsPerWarp + g * kR;
    float s[kR][kKPL];
#pragma unroll
    for (int r = 0; r < kR; ++r)
#pragma unroll
      for (int kk = 0; kk < kKPL; ++kk) s[r][kk] = 0.f;
    for (int kw = 0; kw < words; ++kw) {
      float2 qv[kR];
#pragma unroll
      for (int r = 0; r < kR; ++r) qv[r] = Word<T>::unpack(sQ[(rl + r) * stride + kw]);
#pragma unroll
      for (int kk = 0; kk < kKPL; ++kk) {
        const float2 kv = Word<T>::unpack(sK[jc[kk] * stride + kw]);
#pragma unroll
        for (int r = 0; r < kR; ++r) dot_step<T>(s[r][kk], qv[r], kv);
      }
    }
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int row = row0 + rl + r;
#pragma unroll
      for (int kk = 0; kk < kKPL; ++kk) {
        const int j = lane + 32 * kk;
        if (row < N && j < Tctx) {
          float val = s[r][kk] * scale;
          if (sb.mask != nullptr) val += __ldg(sb.mask + (int64_t)bh * sb.mask_stride_bh + (int64_t)row * sb.mask_stride_n + j);
          const uint32_t flat = (uint32_t)(((int64_t)bh * N + row) * Tctx + j);
          const unsigned long long pk = ((unsigned long long)float_to_ordered(val) << 32) | (0xffffffffu - flat);
          best = pk > best ? pk : best;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if (lane == 0 && best != 0ull) atomicMax(smax, best);
}

// Gradient through the max (paint-with-words): S' = S + mask * coef * M with M = S[argmax], so
// d loss / d S[argmax] += coef * sum(mask o dS').  The backward kernel leaves one partial sum per CTA (fixed order:
// deterministic); this one-CTA kernel adds them up and applies the rank-1 correction to the one query row that owns
// the max:  dQ[b*, n*, h*, :] += scale * gM * K[b*, t*, h*, :].
template <typename T>
__global__ void __launch_bounds__(256)
pww_fixup_kernel(const float* __restrict__ partials, int n_partials, const T* __restrict__ k, T* __restrict__ d_q,
                 int H, int N, int Tctx, int d, float scale, const ga_score_bias_t sb) {
  __shared__ float red[8];
  float t = 0.f;
  for (int i = threadIdx.x; i < n_partials; i += 256) t += partials[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  float total = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) total += red[w];
  const unsigned long long pk = *sb.pww_smax;
  if (pk == 0ull) return;
  const float g_max = __ldg(sb.pww_coef) * total;
  const uint32_t flat = 0xffffffffu - (uint32_t)(pk & 0xffffffffull);
  const int tt = flat % Tctx;
  const int64_t rn = flat / Tctx;
  const int n = (int)(rn % N), bh = (int)(rn / N), b = bh / H, h = bh - b * H;
  const int64_t qoff = (((int64_t)b * N + n) * H + h) * d, koff = (((int64_t)b * Tctx + tt) * H + h) * d;
  for (int c = threadIdx.x; c < d; c += 256)
    d_q[qoff + c] = static_cast<T>(static_cast<float>(d_q[qoff + c]) + scale * g_max * static_cast<float>(k[koff + c]));
}

// ----------------------------------------------------------------------------------------------------- backward
template <typename T, bool kBias>
__global__ void __launch_bounds__(kThreads)
cross_attn_bwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                      const float* __restrict__ lse, const T* __restrict__ d_o, const float* __restrict__ d_acc,
                      int64_t d_acc_bstride, int d_acc_rstride, T* __restrict__ d_q, float* __restrict__ d_k,
                      float* __restrict__ d_v, int H, int N, int Tctx, int d, float scale, const ga_score_bias_t sb,
                      float* __restrict__ pww_partials) {
  constexpr int E = Cfg<T>::E;
  constexpr int kWPL = Cfg<T>::kWPL;
  extern __shared__ uint32_t smem[];
  const int words = d / E;
  const int stride = odd_stride(words);
  uint32_t* sK = smem;
  uint32_t* sV = sK + Tctx * stride;
  uint32_t* sQ = sV + Tctx * stride;
  uint32_t* sG = sQ + kRowsPerCta * stride;  // dO tile
  float* sDS = reinterpret_cast<float*>(sG + kRowsPerCta * stride);
  float* sPP = sDS + kWarps * kR * kPStride;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kRowsPerCta;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int64_t pitch = (int64_t)H * words;
  const int64_t tok_off = ((int64_t)b * N + row0) * pitch + (int64_t)h * words;
  const int64_t ctx_off = (int64_t)b * Tctx * pitch + (int64_t)h * words;

  stage_rows(sK, reinterpret_cast<const uint32_t*>(k) + ctx_off, Tctx, Tctx, words, pitch, stride);
  stage_rows(sV, reinterpret_cast<const uint32_t*>(v) + ctx_off, Tctx, Tctx, words, pitch, stride);
  stage_rows(sQ, reinterpret_cast<const uint32_t*>(q) + tok_off, kRowsPerCta, N - row0, words, pitch, stride);
  stage_rows(sG, reinterpret_cast<const uint32_t*>(d_o) + tok_off, kRowsPerCta, N - row0, words, pitch, stride);
  __syncthreads();

  int jc[kKPL];
#pragma unroll
  for (int kk = 0; kk < kKPL; ++kk) jc[kk] = min(lane + 32 * kk, Tctx - 1);
  BiasLane bl;
  if (kBias) bl = bias_lane_setup(sb, lane);
  float pww_g = 0.f;   // this thread's share of sum(mask o dS') (paint-with-words: gradient through the max)

#pragma unroll 1
  for (int g = 0; g < kGroups; ++g) {
    const int rl = warp * kRowsPerWarp + g * kR;
    float s[kR][kKPL], dp[kR][kKPL];
#pragma unroll
    for (int r = 0; r < kR; ++r)
#pragma unroll
      for (int kk = 0; kk < kKPL; ++kk) s[r][kk] = dp[r][kk] = 0.f;
    for (int kw = 0; kw < words; ++kw) {
      float2 qv[kR], gv[kR];
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        qv[r] = Word<T>::unpack(sQ[(rl + r) * stride + kw]);
        gv[r] = Word<T>::unpack(sG[(rl + r) * stride + kw]);
      }
#pragma unroll
      for (int kk = 0; kk < kKPL; ++kk) {
        const float2 kv = Word<T>::unpack(sK[jc[kk] * stride + kw]);
        const float2 vv = Word<T>::unpack(sV[jc[kk] * stride + kw]);
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          dot_step<T>(s[r][kk], qv[r], kv);
          dot_step<T>(dp[r][kk], gv[r], vv);
        }
      }
    }
    float* dsw = sDS + (warp * kR) * kPStride;
    float* ppw = sPP + (warp * kR) * kPStride;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int row = row0 + rl + r;
      const bool live = row < N;
      const float l = live ? lse[(int64_t)bh * N + row] : 0.f;
      float p[kKPL], dsum = 0.f;
      uint32_t bits = 0u;
      if (kBias) bits = pww_row_bits(sb, row, N, lane);
#pragma unroll
      for (int kk = 0; kk < kKPL; ++kk) {
        const int j = lane + 32 * kk;
        const bool ok = live && j < Tctx;
        float sc = s[r][kk] * scale;
        if (kBias) sc += score_bias(sb, bl, bits, bh, row, N, j, kk, Tctx);
        p[kk] = ok ? expf(sc - l) : 0.f;
        if (ok && d_acc != nullptr) dp[r][kk] += d_acc[(int64_t)b * d_acc_bstride + (int64_t)row * d_acc_rstride + j];
        dsum += p[kk] * dp[r][kk];
      }
      dsum = warp_sum(dsum);
#pragma unroll
      for (int kk = 0; kk < kKPL; ++kk) {
        const int j = lane + 32 * kk;
        const float ds_raw = p[kk] * (dp[r][kk] - dsum);
        if (kBias) pww_g += ds_raw * (float)__popc(bits & bl.colmask[kk]);
        if (j < Tctx) {
          dsw[r * kPStride + j] = ds_raw * scale;
          ppw[r * kPStride + j] = p[kk];
        }
      }
    }
    __syncwarp();

    // dQ = scale * dS K   (scale already folded into dsw)
    float acc[kR][kWPL][E];
#pragma unroll
    for (int r = 0; r < kR; ++r)
#pragma unroll
      for (int m = 0; m < kWPL; ++m)
#pragma unroll
        for (int e2 = 0; e2 < E; ++e2) acc[r][m][e2] = 0.f;
    for (int j = 0; j < Tctx; ++j) {
      float dj[kR];
#pragma unroll
      for (int r = 0; r < kR; ++r) dj[r] = dsw[r * kPStride + j];
#pragma unroll
      for (int m = 0; m < kWPL; ++m) {
        const int cw = lane + 32 * m;
        if (cw < words) {
          const float2 kv = Word<T>::unpack(sK[j * stride + cw]);
#pragma unroll
          for (int r = 0; r < kR; ++r) {
            acc[r][m][0] = fmaf(dj[r], kv.x, acc[r][m][0]);
            if (E == 2) acc[r][m][E - 1] = fmaf(dj[r], kv.y, acc[r][m][E - 1]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int row = row0 + rl + r;
      if (row >= N) continue;
      uint32_t* gq = reinterpret_cast<uint32_t*>(d_q) + ((int64_t)b * N + row) * pitch + (int64_t)h * words;
#pragma unroll
      for (int m = 0; m < kWPL; ++m) {
        const int cw = lane + 32 * m;
        if (cw < words) gq[cw] = Word<T>::pack(acc[r][m][0], acc[r][m][E - 1]);
      }
    }

    // optional text-side gradients (never requested by the guidance path; used by the autograd tests)
    if (d_k != nullptr || d_v != nullptr) {
      for (int j = 0; j < Tctx; ++j) {
        float dj[kR], pj[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          dj[r] = dsw[r * kPStride + j];
          pj[r] = ppw[r * kPStride + j];
        }
#pragma unroll
        for (int m = 0; m < kWPL; ++m) {
          const int cw = lane + 32 * m;
          if (cw >= words) continue;
          float kx = 0.f, ky = 0.f, vx = 0.f, vy = 0.f;
#pragma unroll
          for (int r = 0; r < kR; ++r) {
            const float2 qv = Word<T>::unpack(sQ[(rl + r) * stride + cw]);
            const float2 gv = Word<T>::unpack(sG[(rl + r) * stride + cw]);
            kx = fmaf(dj[r], qv.x, kx); ky = fmaf(dj[r], qv.y, ky);
            vx = fmaf(pj[r], gv.x, vx); vy = fmaf(pj[r], gv.y, vy);
          }
          const int64_t off = ((int64_t)b * Tctx + j) * (int64_t)H * d + (int64_t)h * d + (int64_t)cw * E;
          if (d_k != nullptr) { atomicAdd(d_k + off, kx); if (E == 2) atomicAdd(d_k + off + 1, ky); }
          if (d_v != nullptr) { atomicAdd(d_v + off, vx); if (E == 2) atomicAdd(d_v + off + 1, vy); }
        }
      }
    }
    __syncwarp();
  }
  if (kBias && pww_partials != nullptr) {
    __syncthreads();                       // sDS is free: reuse its first words for the per-warp sums
    pww_g = warp_sum(pww_g);
    if (lane == 0) sDS[warp] = pww_g;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += sDS[w];
      pww_partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------------- launch
// Raise the opt-in dynamic shared-memory limit once per (device, kernel instantiation): no attribute calls afterwards,
// which keeps steady-state launches cheap and CUDA-graph capture clean.
template <typename T>
static cudaError_t ensure_smem(const void* kernel, int slot) {
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

static const ga_score_bias_t kNoBias = {};

template <typename T>
int launch_fwd(const void* q, const void* k, const void* v, void* o, float* lse, float* acc, void* probs, int B, int H,
               int N, int Tctx, int d, float scale, const ga_score_bias_t* bias, cudaStream_t st) {
  const int words = d / Word<T>::E, stride = words | 1;
  const bool probs_only = probs != nullptr;
  size_t smem = (size_t)(2 * Tctx + kRowsPerCta) * stride * 4 + (size_t)kWarps * kR * kPStride * 4;
  if (smem > 226 * 1024) return fail(GA_ERR_UNSUPPORTED, "SIMT cross-attention: head_dim %d needs %zu B smem", d, smem);
  const int heads_per_cta = (acc != nullptr && !probs_only) ? H : 1;
  dim3 grid((N + kRowsPerCta - 1) / kRowsPerCta, B * H / heads_per_cta);
  const bool biased = bias != nullptr;
  auto kern = probs_only ? (biased ? cross_attn_fwd_kernel<T, true, true> : cross_attn_fwd_kernel<T, true, false>)
                         : (biased ? cross_attn_fwd_kernel<T, false, true> : cross_attn_fwd_kernel<T, false, false>);
  cudaError_t e = ensure_smem<T>(reinterpret_cast<const void*>(kern), (probs_only ? 1 : 0) + (biased ? 4 : 0));
  if (e != cudaSuccess) return fail(GA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  kern<<<grid, kThreads, smem, st>>>((const T*)q, (const T*)k, (const T*)v, (T*)o, lse, acc, (T*)probs, H, N, Tctx, d,
                                    scale, heads_per_cta, biased ? *bias : kNoBias);
  return check_launch("cross_attn_fwd_simt");
}

template <typename T>
int launch_bwd(const void* q, const void* k, const void* v, const float* lse, const void* d_o, const float* d_acc,
               int64_t bstride, int rstride, void* d_q, float* d_k, float* d_v, int B, int H, int N, int Tctx, int d, float scale,
               const ga_score_bias_t* bias, float* pww_partials, cudaStream_t st) {
  const int words = d / Word<T>::E, stride = words | 1;
  size_t smem = (size_t)(2 * Tctx + 2 * kRowsPerCta) * stride * 4 + (size_t)2 * kWarps * kR * kPStride * 4;
  if (smem > 226 * 1024) return fail(GA_ERR_UNSUPPORTED, "SIMT cross-attention bwd: head_dim %d needs %zu B smem", d, smem);
  dim3 grid((N + kRowsPerCta - 1) / kRowsPerCta, B * H);
  const bool biased = bias != nullptr;
  auto kern = biased ? cross_attn_bwd_kernel<T, true> : cross_attn_bwd_kernel<T, false>;
  cudaError_t e = ensure_smem<T>(reinterpret_cast<const void*>(kern), biased ? 6 : 2);
  if (e != cudaSuccess) return fail(GA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const bool pww = biased && bias->pww_count > 0;
  kern<<<grid, kThreads, smem, st>>>((const T*)q, (const T*)k, (const T*)v, lse, (const T*)d_o, d_acc, bstride,
                                    rstride, (T*)d_q, d_k, d_v, H, N, Tctx, d, scale, biased ? *bias : kNoBias,
                                    pww ? pww_partials : nullptr);
  int rc = check_launch("cross_attn_bwd_simt");
  if (rc != GA_OK || !pww) return rc;
  // dK of the max term is not produced (nobody differentiates w.r.t. the text side on this path; d_k/d_v callers get
  // the softmax part only, which the header states)
  pww_fixup_kernel<T><<<1, 256, 0, st>>>(pww_partials, (int)(grid.x * grid.y), (const T*)k, (T*)d_q, H, N, Tctx, d,
                                          scale, *bias);
  return check_launch("pww_fixup");
}

template <typename T>
int launch_smax(const void* q, const void* k, unsigned long long* smax, int B, int H, int N, int Tctx, int d,
                float scale, const ga_score_bias_t* bias, cudaStream_t st) {
  const int words = d / Word<T>::E, stride = words | 1;
  const size_t smem = (size_t)(Tctx + kRowsPerCta) * stride * 4;
  if (smem > 226 * 1024) return fail(GA_ERR_UNSUPPORTED, "SIMT score max: head_dim %d needs %zu B smem", d, smem);
  auto kern = cross_attn_smax_kernel<T>;
  cudaError_t e = ensure_smem<T>(reinterpret_cast<const void*>(kern), 3);
  if (e != cudaSuccess) return fail(GA_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  if ((e = cudaMemsetAsync(smax, 0, sizeof(unsigned long long), st)) != cudaSuccess)
    return fail(GA_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  dim3 grid((N + kRowsPerCta - 1) / kRowsPerCta, B * H);
  kern<<<grid, kThreads, smem, st>>>((const T*)q, (const T*)k, smax, H, N, Tctx, d, scale, bias ? *bias : kNoBias);
  return check_launch("cross_attn_smax");
}

int fwd(const void* q, const void* k, const void* v, void* o, float* lse, float* acc, void* probs, int B, int H, int N,
        int Tctx, int d, float scale, int dtype, const ga_score_bias_t* bias, cudaStream_t st) {
  switch (dtype) {
    case GA_F32: return launch_fwd<float>(q, k, v, o, lse, acc, probs, B, H, N, Tctx, d, scale, bias, st);
    case GA_F16: return launch_fwd<__half>(q, k, v, o, lse, acc, probs, B, H, N, Tctx, d, scale, bias, st);
    case GA_BF16: return launch_fwd<__nv_bfloat16>(q, k, v, o, lse, acc, probs, B, H, N, Tctx, d, scale, bias, st);
  }
  return fail(GA_ERR_BAD_ARG, "unknown dtype %d", dtype);
}

int bwd(const void* q, const void* k, const void* v, const float* lse, const void* d_o, const float* d_acc,
        int64_t bstride, int rstride, void* d_q, float* d_k, float* d_v, int B, int H, int N, int Tctx, int d, float scale, int dtype,
        const ga_score_bias_t* bias, float* pww_partials, cudaStream_t st) {
  switch (dtype) {
    case GA_F32: return launch_bwd<float>(q, k, v, lse, d_o, d_acc, bstride, rstride, d_q, d_k, d_v, B, H, N, Tctx, d, scale, bias, pww_partials, st);
    case GA_F16: return launch_bwd<__half>(q, k, v, lse, d_o, d_acc, bstride, rstride, d_q, d_k, d_v, B, H, N, Tctx, d, scale, bias, pww_partials, st);
    case GA_BF16:
      return launch_bwd<__nv_bfloat16>(q, k, v, lse, d_o, d_acc, bstride, rstride, d_q, d_k, d_v, B, H, N, Tctx, d, scale, bias, pww_partials, st);
  }
  return fail(GA_ERR_BAD_ARG, "unknown dtype %d", dtype);
}

int smax(const void* q, const void* k, unsigned long long* out, int B, int H, int N, int Tctx, int d, float scale,
         int dtype, const ga_score_bias_t* bias, cudaStream_t st) {
  switch (dtype) {
    case GA_F32: return launch_smax<float>(q, k, out, B, H, N, Tctx, d, scale, bias, st);
    case GA_F16: return launch_smax<__half>(q, k, out, B, H, N, Tctx, d, scale, bias, st);
    case GA_BF16: return launch_smax<__nv_bfloat16>(q, k, out, B, H, N, Tctx, d, scale, bias, st);
  }
  return fail(GA_ERR_BAD_ARG, "unknown dtype %d", dtype);
}

}  // namespace simt
}  // namespace ga
