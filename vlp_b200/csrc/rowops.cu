// vlp_b200 — HBM-bound row kernels of the VLP hot path (warp-per-row, 16-byte vector access,
// fp32 statistics through __shfl_xor_sync).  No tensor cores here by design: these are pure
// bandwidth kernels and their roofline is HBM GB/s.
//
//   ln_res_drop_fwd/bwd : y = LN(dropout(t) + res)           BertSelfOutput / BertOutput, modeling.py:313-317, 353-357
//   embed_fwd/bwd       : y = dropout(LN(word|vis + pos|vis_pe + type))      BertEmbeddings, modeling.py:217-241
//   mask_pack           : additive/0-1 attention mask -> 128-bit row bitmask  get_extended_attention_mask, modeling.py:807-833
//   colsum              : bias gradients
//   f32_to_bf16         : gradient arena conversion
#include "rowops.cuh"
#include "host.cuh"

namespace vlpk {

static constexpr int MAXCH = 4;  // up to 4 x (32 lanes x 8 elems) = 1024 columns per row

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = unpack_bf16x2(w[j]);
    f[2 * j] = x.x;
    f[2 * j + 1] = x.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// mean / rstd of a row distributed as z[i][0..7] over the warp (two-pass, fp32)
__device__ __forceinline__ void row_stats(const float (&z)[MAXCH][8], const bool (&ok)[MAXCH], int H, float eps,
                                          float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (ok[i]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s += z[i][j];
    }
  mean = warp_sum(s) / H;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (ok[i]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = z[i][j] - mean;
        v += d * d;
      }
    }
  rstd = rsqrtf(warp_sum(v) / H + eps);
}

// ------------------------------------------------------------------------------------------------
// LN + residual + dropout
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_res_drop_fwd_kernel(LnArgs a) {
  const uint64_t dseed = drop_seed(a.drop);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long rowi = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5);
  if (rowi >= a.M) return;
  float z[MAXCH][8];
  bool ok[MAXCH];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int col = (lane + 32 * i) * 8;
    ok[i] = col < a.H;
    if (ok[i]) {
      float t[8];
      load8(a.t + rowi * a.H + col, t);
      uint32_t keep = 0xFFu;
      if (a.drop.p > 0.f) keep = dropout_keep8(dseed, a.drop.site, (rowi * a.H + col) >> 3, a.drop.thresh16);
      float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (a.res != nullptr) load8(a.res + rowi * a.H + col, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[i][j] = (((keep >> j) & 1u) ? t[j] * a.drop.scale : 0.f) + r[j];
    }
  }
  float mean, rstd;
  row_stats(z, ok, a.H, a.eps, mean, rstd);
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (ok[i]) {
      float g[8], be[8], y[8];
      load8(a.gamma + col, g);
      load8(a.beta + col, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (z[i][j] - mean) * rstd * g[j] + be[j];
      store8(a.y + rowi * a.H + col, y);
    }
  }
  if (lane == 0 && a.stats != nullptr) a.stats[rowi] = make_float2(mean, rstd);
}

// Backward.  Persistent warps accumulate dgamma/dbeta(/dbias) over their rows in registers, then one
// shared-memory reduction + one fp32 atomicAdd per column per block.
__global__ void __launch_bounds__(256) ln_res_drop_bwd_kernel(LnArgs a) {
  const uint64_t dseed = drop_seed(a.drop);
  extern __shared__ float s_red[];  // [wpb][3][H]
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  float ag[MAXCH][8], ab[MAXCH][8], at[MAXCH][8];
  bool ok[MAXCH];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    ok[i] = (lane + 32 * i) * 8 < a.H;
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = at[i][j] = 0.f;
  }
  float g[MAXCH][8];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (ok[i]) load8(a.gamma + (lane + 32 * i) * 8, g[i]);

  for (long long rowi = static_cast<long long>(blockIdx.x) * wpb + wib; rowi < a.M;
       rowi += static_cast<long long>(gridDim.x) * wpb) {
    const float2 st = a.stats[rowi];
    const float mean = st.x, rstd = st.y;
    float xh[MAXCH][8], gy[MAXCH][8];
    uint32_t keepm[MAXCH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int col = (lane + 32 * i) * 8;
      keepm[i] = 0xFFu;
      if (ok[i]) {
        float t[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dy[8];
        load8(a.t + rowi * a.H + col, t);
        if (a.res != nullptr) load8(a.res + rowi * a.H + col, r);
        load8(a.dy + rowi * a.H + col, dy);
        if (a.drop.p > 0.f) keepm[i] = dropout_keep8(dseed, a.drop.site, (rowi * a.H + col) >> 3, a.drop.thresh16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float zz = (((keepm[i] >> j) & 1u) ? t[j] * a.drop.scale : 0.f) + r[j];
          xh[i][j] = (zz - mean) * rstd;
          gy[i][j] = dy[j] * g[i][j];
          s1 += gy[i][j];
          s2 += gy[i][j] * xh[i][j];
          ag[i][j] += dy[j] * xh[i][j];
          ab[i][j] += dy[j];
        }
      }
    }
    const float c1 = warp_sum(s1) / a.H, c2 = warp_sum(s2) / a.H;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int col = (lane + 32 * i) * 8;
      if (ok[i]) {
        float dz[8], dt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dz[j] = rstd * (gy[i][j] - c1 - xh[i][j] * c2);
          dt[j] = ((keepm[i] >> j) & 1u) ? dz[j] * a.drop.scale : 0.f;
          at[i][j] += dt[j];
        }
        if (a.dz != nullptr) store8(a.dz + rowi * a.H + col, dz);
        if (a.dt != nullptr) store8(a.dt + rowi * a.H + col, dt);
      }
    }
  }
  // block reduction
  float* sg = s_red + (wib * 3 + 0) * a.H;
  float* sb = s_red + (wib * 3 + 1) * a.H;
  float* stt = s_red + (wib * 3 + 2) * a.H;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (ok[i]) {
      const int col = (lane + 32 * i) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sg[col + j] = ag[i][j];
        sb[col + j] = ab[i][j];
        stt[col + j] = at[i][j];
      }
    }
  __syncthreads();
  for (int c = threadIdx.x; c < a.H; c += blockDim.x) {
    float vg = 0.f, vb = 0.f, vt = 0.f;
    for (int w = 0; w < wpb; ++w) {
      vg += s_red[(w * 3 + 0) * a.H + c];
      vb += s_red[(w * 3 + 1) * a.H + c];
      vt += s_red[(w * 3 + 2) * a.H + c];
    }
    if (a.dgamma != nullptr) atomicAdd(a.dgamma + c, vg);
    if (a.dbeta != nullptr) atomicAdd(a.dbeta + c, vb);
    if (a.dbias != nullptr) atomicAdd(a.dbias + c, vt);
  }
}

static int check_ln(const LnArgs& a) {
  VLPK_CHECK_ARG(a.M > 0 && a.H > 0 && a.H % 8 == 0 && a.H <= MAXCH * 256, "layernorm: H=%d must be a multiple of 8 and <= %d",
                 a.H, MAXCH * 256);
  return 0;
}

int launch_ln_res_drop_fwd(const LnArgs& a, cudaStream_t s) {
  VLPK_TRY(check_ln(a));
  const int wpb = 8;
  const long long grid = (a.M + wpb - 1) / wpb;
  LaunchScope scope(CAT_LN_FWD, 2.0 * a.M * a.H * (a.res ? 3 : 2) + 8.0 * a.M, s);
  ln_res_drop_fwd_kernel<<<static_cast<unsigned>(grid), wpb * 32, 0, s>>>(a);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

int launch_ln_res_drop_bwd(const LnArgs& a, cudaStream_t s) {
  VLPK_TRY(check_ln(a));
  const int wpb = 8;
  long long grid = (a.M + wpb - 1) / wpb;
  const long long cap = static_cast<long long>(num_sms()) * 4;
  if (grid > cap) grid = cap;
  const size_t smem = static_cast<size_t>(wpb) * 3 * a.H * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    VLPK_CUDA(cudaFuncSetAttribute(ln_res_drop_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 1024 * 4));
    attr_set = true;
  }
  LaunchScope scope(CAT_LN_BWD, 2.0 * a.M * a.H * ((a.res ? 3 : 2) + (a.dz ? 1 : 0) + (a.dt ? 1 : 0)) + 8.0 * a.M, s);
  ln_res_drop_bwd_kernel<<<static_cast<unsigned>(grid), wpb * 32, smem, s>>>(a);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Embeddings: gather + region splice + LN + dropout
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void embed_row_z(const EmbedArgs& a, long long rowi, int lane, float (&z)[MAXCH][8],
                                            bool (&ok)[MAXCH]) {
  const int b = static_cast<int>(rowi / a.L), l = static_cast<int>(rowi % a.L);
  const bool vis = a.vis_input && l >= 1 && l <= a.R;
  const long long wid = a.ids[rowi];
  const long long pid = a.pos != nullptr ? a.pos[rowi] : l;
  const long long tid = a.tt != nullptr ? a.tt[rowi] : 0;
  const __nv_bfloat16* wsrc = vis ? a.vis + (static_cast<long long>(b) * a.R + (l - 1)) * a.H : a.word + wid * a.H;
  const __nv_bfloat16* psrc = vis ? a.vpe + (static_cast<long long>(b) * a.R + (l - 1)) * a.H : a.posw + pid * a.H;
  const __nv_bfloat16* tsrc = a.typew + tid * a.H;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int col = (lane + 32 * i) * 8;
    ok[i] = col < a.H;
    if (ok[i]) {
      float w[8], p[8], t[8];
      load8(wsrc + col, w);
      load8(psrc + col, p);
      load8(tsrc + col, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[i][j] = w[j] + p[j] + t[j];
    }
  }
}

__global__ void __launch_bounds__(256) embed_fwd_kernel(EmbedArgs a) {
  const uint64_t dseed = drop_seed(a.drop);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long rowi = static_cast<long long>(blockIdx.x) * wpb + (threadIdx.x >> 5);
  const long long M = static_cast<long long>(a.B) * a.L;
  if (rowi >= M) return;
  float z[MAXCH][8];
  bool ok[MAXCH];
  embed_row_z(a, rowi, lane, z, ok);
  float mean, rstd;
  row_stats(z, ok, a.H, a.eps, mean, rstd);
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int col = (lane + 32 * i) * 8;
    if (ok[i]) {
      float g[8], be[8], y[8];
      load8(a.gamma + col, g);
      load8(a.beta + col, be);
      uint32_t keep = 0xFFu;
      if (a.drop.p > 0.f) keep = dropout_keep8(dseed, a.drop.site, (rowi * a.H + col) >> 3, a.drop.thresh16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = (z[i][j] - mean) * rstd * g[j] + be[j];
        y[j] = ((keep >> j) & 1u) ? v * a.drop.scale : 0.f;
      }
      store8(a.y + rowi * a.H + col, y);
    }
  }
  if (lane == 0 && a.stats != nullptr) a.stats[rowi] = make_float2(mean, rstd);
}

__global__ void __launch_bounds__(256) embed_bwd_kernel(EmbedArgs a) {
  const uint64_t dseed = drop_seed(a.drop);
  extern __shared__ float s_red[];  // [wpb][2][H]
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const long long M = static_cast<long long>(a.B) * a.L;
  float ag[MAXCH][8], ab[MAXCH][8], g[MAXCH][8];
  bool okc[MAXCH];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    okc[i] = (lane + 32 * i) * 8 < a.H;
    if (okc[i]) load8(a.gamma + (lane + 32 * i) * 8, g[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = 0.f;
  }
  for (long long rowi = static_cast<long long>(blockIdx.x) * wpb + wib; rowi < M;
       rowi += static_cast<long long>(gridDim.x) * wpb) {
    float z[MAXCH][8];
    bool ok[MAXCH];
    embed_row_z(a, rowi, lane, z, ok);
    const float2 st = a.stats[rowi];
    const float mean = st.x, rstd = st.y;
    float gy[MAXCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int col = (lane + 32 * i) * 8;
      if (ok[i]) {
        float dy[8];
        load8(a.dy + rowi * a.H + col, dy);
        uint32_t keep = 0xFFu;
        if (a.drop.p > 0.f) keep = dropout_keep8(dseed, a.drop.site, (rowi * a.H + col) >> 3, a.drop.thresh16);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = ((keep >> j) & 1u) ? dy[j] * a.drop.scale : 0.f;  // grad wrt LN output
          z[i][j] = (z[i][j] - mean) * rstd;                                // x-hat
          gy[i][j] = d * g[i][j];
          s1 += gy[i][j];
          s2 += gy[i][j] * z[i][j];
          ag[i][j] += d * z[i][j];
          ab[i][j] += d;
        }
      }
    }
    const float c1 = warp_sum(s1) / a.H, c2 = warp_sum(s2) / a.H;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int col = (lane + 32 * i) * 8;
      if (ok[i]) {
        float dz[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = rstd * (gy[i][j] - c1 - z[i][j] * c2);
        store8(a.dz + rowi * a.H + col, dz);
      }
    }
  }
  float* sg = s_red + (wib * 2 + 0) * a.H;
  float* sb = s_red + (wib * 2 + 1) * a.H;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (okc[i]) {
      const int col = (lane + 32 * i) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sg[col + j] = ag[i][j];
        sb[col + j] = ab[i][j];
      }
    }
  __syncthreads();
  for (int c = threadIdx.x; c < a.H; c += blockDim.x) {
    float vg = 0.f, vb = 0.f;
    for (int w = 0; w < wpb; ++w) {
      vg += s_red[(w * 2 + 0) * a.H + c];
      vb += s_red[(w * 2 + 1) * a.H + c];
    }
    atomicAdd(a.dgamma + c, vg);
    atomicAdd(a.dbeta + c, vb);
  }
}

static int check_embed(const EmbedArgs& a) {
  VLPK_CHECK_ARG(a.B > 0 && a.L > 0 && a.H > 0 && a.H % 8 == 0 && a.H <= MAXCH * 256, "embed: bad shape B=%d L=%d H=%d", a.B,
                 a.L, a.H);
  VLPK_CHECK_ARG(!a.vis_input || (a.vis != nullptr && a.vpe != nullptr && a.R + 1 <= a.L), "embed: vis_input needs vis/vpe and R+1<=L");
  return 0;
}

int launch_embed_fwd(const EmbedArgs& a, cudaStream_t s) {
  VLPK_TRY(check_embed(a));
  const int wpb = 8;
  const long long M = static_cast<long long>(a.B) * a.L;
  LaunchScope scope(CAT_EMBED, 2.0 * M * a.H * 4, s);
  embed_fwd_kernel<<<static_cast<unsigned>((M + wpb - 1) / wpb), wpb * 32, 0, s>>>(a);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

int launch_embed_bwd(const EmbedArgs& a, cudaStream_t s) {
  VLPK_TRY(check_embed(a));
  VLPK_CHECK_ARG(a.dy && a.dz && a.stats && a.dgamma && a.dbeta, "embed bwd: missing buffers");
  const int wpb = 8;
  const long long M = static_cast<long long>(a.B) * a.L;
  long long grid = (M + wpb - 1) / wpb;
  const long long cap = static_cast<long long>(num_sms()) * 4;
  if (grid > cap) grid = cap;
  const size_t smem = static_cast<size_t>(wpb) * 2 * a.H * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    VLPK_CUDA(cudaFuncSetAttribute(embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4));
    attr_set = true;
  }
  LaunchScope scope(CAT_EMBED, 2.0 * M * a.H * 5, s);
  embed_bwd_kernel<<<static_cast<unsigned>(grid), wpb * 32, smem, s>>>(a);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// mask pack
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool mask_attend(T v, int mode);
template <>
__device__ __forceinline__ bool mask_attend<float>(float v, int mode) { return mode == MASK_ADDITIVE ? v > -5000.f : v != 0.f; }
template <>
__device__ __forceinline__ bool mask_attend<__nv_bfloat16>(__nv_bfloat16 v, int mode) {
  return mask_attend<float>(__bfloat162float(v), mode);
}
template <>
__device__ __forceinline__ bool mask_attend<long long>(long long v, int mode) { return v != 0; }

template <typename T>
__global__ void mask_pack_kernel(const T* m, long long sb, long long sr, int B, int rows, int kv, int mode, uint32_t* out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * rows) return;
  const int b = static_cast<int>(idx / rows), r = static_cast<int>(idx % rows);
  const T* p = m + b * sb + r * sr;
  uint32_t w[4] = {0, 0, 0, 0};
  for (int j = 0; j < kv; ++j)
    if (mask_attend<T>(p[j], mode)) w[j >> 5] |= 1u << (j & 31);
  reinterpret_cast<uint4*>(out)[idx] = make_uint4(w[0], w[1], w[2], w[3]);
}

int launch_mask_pack(const void* mask, int dtype, int mode, int B, int rows, int kv, long long stride_b, long long stride_r,
                     uint32_t* out, cudaStream_t s) {
  VLPK_CHECK_ARG(B > 0 && rows > 0 && kv > 0 && kv <= 128, "mask_pack: kv=%d must be in [1,128]", kv);
  const long long n = static_cast<long long>(B) * rows;
  const unsigned grid = static_cast<unsigned>((n + 127) / 128);
  LaunchScope scope(CAT_MISC, 0.0, s);
  switch (dtype) {
    case VLPK_DT_F32: mask_pack_kernel<float><<<grid, 128, 0, s>>>(static_cast<const float*>(mask), stride_b, stride_r, B, rows, kv, mode, out); break;
    case VLPK_DT_BF16: mask_pack_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(mask), stride_b, stride_r, B, rows, kv, mode, out); break;
    case VLPK_DT_I64: mask_pack_kernel<long long><<<grid, 128, 0, s>>>(static_cast<const long long*>(mask), stride_b, stride_r, B, rows, kv, mode, out); break;
    default: set_error("mask_pack: unsupported dtype %d", dtype); return -1;
  }
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients) and fp32 -> bf16 conversion
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* x, long long ld, long long M, int N, int rows_per_blk,
                                                      float* out) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (col >= N) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_blk;
  const long long r1 = min(M, r0 + rows_per_blk);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = r0; r < r1; ++r) {
    float v[8];
    load8(x + r * ld + col, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(out + col + j, acc[j]);
}

int launch_colsum(const void* x, long long ld, long long M, int N, float* out, cudaStream_t s) {
  VLPK_CHECK_ARG(N % 8 == 0 && ld % 8 == 0, "colsum: N=%d ld=%lld must be multiples of 8", N, ld);
  const int rows_per_blk = 32;
  dim3 grid((N / 8 + 255) / 256, static_cast<unsigned>((M + rows_per_blk - 1) / rows_per_blk));
  LaunchScope scope(CAT_MISC, 2.0 * M * N, s);
  colsum_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ld, M, N, rows_per_blk, out);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + i + 4));
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y);
    u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y);
    u.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(y + i) = u;
  } else {
    for (long long k = i; k < n; ++k) y[k] = __float2bfloat16_rn(x[k]);
  }
}

int launch_f32_to_bf16(const float* x, void* y, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  VLPK_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0,
                 "f32_to_bf16: pointers must be 16-byte aligned");
  const long long nthreads = (n + 7) / 8;
  LaunchScope scope(CAT_MISC, 6.0 * n, s);
  f32_to_bf16_kernel<<<static_cast<unsigned>((nthreads + 255) / 256), 256, 0, s>>>(x, static_cast<__nv_bfloat16*>(y), n);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vlpk
