// vlp_b200 — scatter of BertEmbeddings' pre-LayerNorm gradient into the word / position / token-type tables
// (autograd backward of the three nn.Embedding lookups, modeling.py:217-241).
//
// In the round-1 profile this glue, done with torch ops, cost ~215 us of a 7.2 ms step: a dense fp32 zero-fill of the
// [28996,768] word table (89 MB), its conversion to bf16, and an index_add of all B*L rows into the 6-row token-type table
// (7 872 x 768 atomics onto 6 x 768 addresses).  Only B x 23 rows of the word / position tables are ever looked up
// (the 100 region rows are spliced in from the projections), so:
//   word : memset the bf16 gradient (44 MB), then three tiny launches over the looked-up rows only — zero their fp32
//          scratch rows, atomically accumulate (duplicates such as [CLS]/[SEP] collide here, in fp32), convert to bf16;
//   pos  : fp32 atomics from the same rows;
//   type : a segmented column sum — each thread keeps one accumulator per token type (<= 8) — one atomic per
//          (type, column, 256-row slab) instead of one per element.
#include "tables.cuh"

#include "host.cuh"

namespace vlpk {
namespace {

constexpr int TT_MAX = 8;
constexpr int SLAB = 256;

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16x2(w[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
}

// entry e -> (sample, position) of the e-th row that reads the word / position tables
__device__ __forceinline__ long long table_row(const TableGradArgs& a, long long e) {
  const int n_tab = a.vis_input ? a.L - a.R : a.L;
  const long long b = e / n_tab;
  const int k = static_cast<int>(e % n_tab);
  const int l = a.vis_input ? (k == 0 ? 0 : a.R + k) : k;
  return b * a.L + l;
}

// PHASE 0: zero the scratch rows; 1: accumulate; 2: scratch -> bf16.  One warp per looked-up row.
template <int PHASE>
__global__ void __launch_bounds__(256) word_pos_kernel(TableGradArgs a, long long n_entries) {
  const long long e = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (e >= n_entries) return;
  const int lane = threadIdx.x & 31;
  const long long row = table_row(a, e);
  const long long id = a.ids[row];
  const bool id_ok = (id >= 0 && id < a.V);
  long long p = (a.pos != nullptr) ? a.pos[row] : (row % a.L);
  const bool p_ok = (p >= 0 && p < a.P);
  for (int c = lane * 8; c < a.H; c += 256) {
    if (PHASE == 0) {
      if (id_ok) {
        float* d = a.scratch + id * a.H + c;
        *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(d + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else if (PHASE == 1) {
      float v[8];
      ld8(a.dz + row * a.H + c, v);
      if (id_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(a.scratch + id * a.H + c + j, v[j]);
      }
      if (p_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(a.d_pos + p * a.H + c + j, v[j]);
      }
    } else {
      if (id_ok) {
        const float* sp = a.scratch + id * a.H + c;
        const float4 x = *reinterpret_cast<const float4*>(sp), y = *reinterpret_cast<const float4*>(sp + 4);
        *reinterpret_cast<uint4*>(a.d_word + id * a.H + c) =
            make_uint4(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w), pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
      }
    }
  }
}

// d_type[t, :] += sum over the rows of this 256-row slab whose token type is t.  Block: 8 column groups x 32 row lanes.
__global__ void __launch_bounds__(256) type_grad_kernel(TableGradArgs a) {
  __shared__ float s_part[32][65];
  const int cgp = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int col = blockIdx.x * 64 + cgp * 8;
  const long long M = static_cast<long long>(a.B) * a.L;
  const long long r0 = static_cast<long long>(blockIdx.y) * SLAB;
  float acc[TT_MAX][8];
#pragma unroll
  for (int t = 0; t < TT_MAX; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  if (col < a.H) {
#pragma unroll
    for (int i = 0; i < SLAB / 32; ++i) {
      const long long r = r0 + rl + 32 * i;
      if (r < M) {
        float v[8];
        ld8(a.dz + r * a.H + col, v);
        const int ty = (a.tt != nullptr) ? static_cast<int>(a.tt[r]) : 0;
#pragma unroll
        for (int t = 0; t < TT_MAX; ++t) {
          const float sel = (ty == t) ? 1.f : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(sel, v[j], acc[t][j]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TT_MAX; ++t) {
    if (t < a.T) {  // uniform across the block
#pragma unroll
      for (int j = 0; j < 8; ++j) s_part[rl][cgp * 8 + j] = acc[t][j];
      __syncthreads();
      if (threadIdx.x < 64) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) tot += s_part[i][threadIdx.x];
        const int c = blockIdx.x * 64 + threadIdx.x;
        if (c < a.H && tot != 0.f) atomicAdd(a.d_type + static_cast<long long>(t) * a.H + c, tot);
      }
      __syncthreads();
    }
  }
}

// Explicit row lists (data parallelism, vlp_b200/dp.py): the looked-up rows of ALL ranks (all-gathered: 23 rows per sample instead of
// all-reducing the dense [V,H] table gradient) are added, scaled by 1/world, INTO an existing bf16 word-table gradient — the tied
// decoder weight's gradient, whose own all-reduce was issued as soon as the head's backward produced it — and into the fp32 position
// gradient.  Duplicate ids (every [CLS] / [SEP]) are summed in fp32 first; one warp per id (elected through `owner`) does the bf16 update.
template <int PHASE>
__global__ void __launch_bounds__(256) table_rows_kernel(TableRowsArgs a) {
  const long long e = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (e >= a.n) return;
  const int lane = threadIdx.x & 31;
  const long long id = a.ids[e];
  const bool id_ok = (id >= 0 && id < a.V);
  const long long p = (a.pos != nullptr) ? a.pos[e] : -1;
  const bool p_ok = (a.d_pos != nullptr && p >= 0 && p < a.P);
  bool won = false;
  if (PHASE == 0) {
    if (id_ok && lane == 0) a.owner[id] = 0;
  } else if (PHASE == 2) {
    int w = 0;
    if (id_ok && lane == 0) w = (atomicExch(a.owner + id, 1) == 0) ? 1 : 0;
    won = __shfl_sync(0xffffffffu, w, 0) != 0;
    if (!won) return;
  }
  for (int c = lane * 8; c < a.H; c += 256) {
    if (PHASE == 0) {
      if (id_ok) {
        float* d = a.scratch + id * a.H + c;
        *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(d + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else if (PHASE == 1) {
      float v[8];
      ld8(a.rows + e * a.H + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= a.scale;
      if (id_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(a.scratch + id * a.H + c + j, v[j]);
      }
      if (p_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(a.d_pos + p * a.H + c + j, v[j]);
      }
    } else {
      float cur[8];
      ld8(a.d_word + id * a.H + c, cur);
      const float* sp = a.scratch + id * a.H + c;
      const float4 x = *reinterpret_cast<const float4*>(sp), y = *reinterpret_cast<const float4*>(sp + 4);
      *reinterpret_cast<uint4*>(a.d_word + id * a.H + c) =
          make_uint4(pack_bf16x2(cur[0] + x.x, cur[1] + x.y), pack_bf16x2(cur[2] + x.z, cur[3] + x.w), pack_bf16x2(cur[4] + y.x, cur[5] + y.y),
                     pack_bf16x2(cur[6] + y.z, cur[7] + y.w));
    }
  }
}

}  // namespace

int launch_table_rows_add(const TableRowsArgs& a, cudaStream_t s) {
  VLPK_CHECK_ARG(a.n > 0 && a.H > 0 && a.H % 8 == 0 && a.V > 0, "table_rows_add: n=%lld H=%d V=%d", a.n, a.H, a.V);
  VLPK_CHECK_ARG(a.ids && a.rows && a.d_word && a.scratch && a.owner, "table_rows_add: null pointer");
  VLPK_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.rows) | reinterpret_cast<uintptr_t>(a.d_word) | reinterpret_cast<uintptr_t>(a.scratch)) & 15u) == 0,
                 "table_rows_add: rows / d_word / scratch must be 16-byte aligned");
  const unsigned grid = static_cast<unsigned>((a.n + 7) / 8);
  {
    LaunchScope scope(CAT_EMBED, 0.0, s);
    table_rows_kernel<0><<<grid, 256, 0, s>>>(a);
    VLPK_CUDA(cudaGetLastError());
  }
  {
    LaunchScope scope(CAT_EMBED, 0.0, s);
    table_rows_kernel<1><<<grid, 256, 0, s>>>(a);
    VLPK_CUDA(cudaGetLastError());
  }
  LaunchScope scope(CAT_EMBED, 0.0, s);
  table_rows_kernel<2><<<grid, 256, 0, s>>>(a);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

int launch_embed_tables_bwd(const TableGradArgs& a, cudaStream_t s) {
  VLPK_CHECK_ARG(a.B > 0 && a.L > 0 && a.H > 0 && a.H % 8 == 0, "embed_tables_bwd: B=%d L=%d H=%d (H must be a multiple of 8)", a.B, a.L, a.H);
  VLPK_CHECK_ARG(a.V > 0 && a.P > 0 && a.T > 0 && a.T <= TT_MAX, "embed_tables_bwd: V=%d P=%d T=%d (at most %d token types)", a.V, a.P, a.T, TT_MAX);
  VLPK_CHECK_ARG(!a.vis_input || (a.R > 0 && a.R < a.L), "embed_tables_bwd: R=%d regions do not fit L=%d", a.R, a.L);
  const bool type_only = (a.d_word == nullptr && a.scratch == nullptr && a.d_pos == nullptr);   // word / position rows handled elsewhere
  VLPK_CHECK_ARG(a.ids && a.dz && a.d_type && (type_only || (a.d_word && a.scratch && a.d_pos)), "embed_tables_bwd: null pointer");
  VLPK_CHECK_ARG(((reinterpret_cast<uintptr_t>(a.dz) | reinterpret_cast<uintptr_t>(a.d_word) | reinterpret_cast<uintptr_t>(a.scratch)) & 15u) == 0,
                 "embed_tables_bwd: dz / d_word / scratch must be 16-byte aligned");
  const long long M = static_cast<long long>(a.B) * a.L;
  const long long n_entries = static_cast<long long>(a.B) * (a.vis_input ? a.L - a.R : a.L);
  const unsigned grid = static_cast<unsigned>((n_entries + 7) / 8);
  if (!type_only) {
    VLPK_CUDA(cudaMemsetAsync(a.d_word, 0, static_cast<size_t>(a.V) * a.H * 2, s));
  }
  if (!type_only) {
    LaunchScope scope(CAT_EMBED, 0.0, s);
    word_pos_kernel<0><<<grid, 256, 0, s>>>(a, n_entries);
    VLPK_CUDA(cudaGetLastError());
  }
  if (!type_only) {
    LaunchScope scope(CAT_EMBED, 0.0, s);
    word_pos_kernel<1><<<grid, 256, 0, s>>>(a, n_entries);
    VLPK_CUDA(cudaGetLastError());
  }
  if (!type_only) {
    LaunchScope scope(CAT_EMBED, 2.0 * a.V * a.H, s);
    word_pos_kernel<2><<<grid, 256, 0, s>>>(a, n_entries);
    VLPK_CUDA(cudaGetLastError());
  }
  LaunchScope scope(CAT_EMBED, 2.0 * M * a.H, s);
  type_grad_kernel<<<dim3((a.H + 63) / 64, static_cast<unsigned>((M + SLAB - 1) / SLAB)), 256, 0, s>>>(a);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vlpk
