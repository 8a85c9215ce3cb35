// vlp_b200 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )         bf16 operands, fp32 accumulation in TMEM
//
// Every dense contraction on the VLP hot path goes through this one kernel family
// (SURVEY.md §8a a1,a2,a5,a6,a8,a9,a18 / Appendix B):
//   forward  Linear      : A = activations [M,K] (K-major),  B = weight [N,K] (K-major)
//   dgrad    dX = dY W   : A = dY [M,N'] (K-major),          B = weight [N',K'] read MN-major (no transpose copy)
//   wgrad    dW = dY^T X : A = dY read MN-major, B = X read MN-major, split-K, fp32 TMA reduce-add
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor -> 128B-swizzled smem ring (STAGES deep)
//   warp 1      MMA issuer     : tcgen05.mma (M=128, N=BN, K=16) x4 per 64-wide k-block, accumulators
//                                double-buffered in TMEM (2 x BN columns) so epilogue(i) overlaps mainloop(i+1)
//   warps 2..5  epilogue       : tcgen05.ld -> registers -> fused pointwise op -> swizzled smem staging
//                                -> TMA store (or TMA reduce-add for split-K weight gradients)
#include "gemm.cuh"
#include "host.cuh"

namespace vlpk {

static constexpr int BM = 128;  // tile M == UMMA M
static constexpr int BK = 64;   // k-block: 64 bf16 = one 128-byte swizzle span
static constexpr int NUM_THREADS = 192;
static constexpr int STG_BYTES = BM * 128;  // one staging buffer: 128 rows x 128 bytes

struct GemmTmaps {
  CUtensorMap a;
  CUtensorMap b[3];
  CUtensorMap d0;
  CUtensorMap d1;
};

struct GemmArgs {
  int M, N, K;
  int b_seg_rows;
  int splits;
  const __nv_bfloat16* bias[3];
  const __nv_bfloat16* aux;
  long long ld_aux;
  float relu_scale;
  DropoutCfg drop;
  // MN-major operand descriptor geometry (bytes).  Fixed by the TMA box layout (see MMA issuer);
  // runtime values only so that bring-up tests can probe the encoding (vlpk_debug_set_mn_desc).
  uint32_t mn_lbo, mn_sbo, mn_kstep;
};

static uint32_t g_mn_lbo = 8192, g_mn_sbo = 1024, g_mn_kstep = 2048;
void debug_set_mn_desc(uint32_t lbo, uint32_t sbo, uint32_t kstep) {
  g_mn_lbo = lbo;
  g_mn_sbo = sbo;
  g_mn_kstep = kstep;
}

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_STG = STAGES * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_STG + 2 * STG_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + BN * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-byte alignment
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Write 8 packed bf16 pairs... (16 bytes) of this thread's row into a 128B-swizzled staging tile.
__device__ __forceinline__ void stg_write16(uint8_t* stg, int row, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(stg + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}

template <int BN, bool A_MN, bool B_MN, int EPI, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ GemmTmaps tm, const GemmArgs args) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);
  float* s_bias = reinterpret_cast<float*>(smem + L::OFF_BIAS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (args.M + BM - 1) / BM;
  const int num_n = (args.N + BN - 1) / BN;
  const int total_kb = (args.K + BK - 1) / BK;
  const int kb_per = (total_kb + args.splits - 1) / args.splits;
  const int num_work = num_m * num_n * args.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.b[0]);
    tma_prefetch_desc(&tm.d0);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer ========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int n_blk = w % num_n;
        const int m_blk = (w / num_n) % num_m;
        const int split = w / (num_n * num_m);
        const int kb0 = split * kb_per;
        const int kb1 = min(total_kb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sA = smem + stage * L::STAGE_BYTES;
          uint8_t* sB = sA + L::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sA, &tm.a, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sA + j * 8192, &tm.a, &full_bar[stage], m_blk * BM + j * 64, kb * BK);
          }
          if (!B_MN) {
            const int n0 = n_blk * BN;
            const int seg = n0 / args.b_seg_rows;
            tma_load_2d(sB, &tm.b[seg], &full_bar[stage], kb * BK, n0 - seg * args.b_seg_rows);
          } else {
            const int k0 = kb * BK;
            const int seg = k0 / args.b_seg_rows;
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sB + j * 8192, &tm.b[seg], &full_bar[stage], n_blk * BN + j * 64,
                          k0 - seg * args.b_seg_rows);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int split = w / (num_n * num_m);
        const int kb0 = split * kb_per;
        const int kb1 = min(total_kb, kb0 + kb_per);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sB = sA + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: step 16 elements (32 B) inside the swizzle span.  SBO = 8 rows x 128 B.
            // MN-major: step 16 k-rows (16 x 128 B).  LBO = next 64-wide MN block, SBO = next 8 k-rows.
            const uint64_t adesc = A_MN ? umma_smem_desc_sw128(sA + k * args.mn_kstep, args.mn_lbo, args.mn_sbo)
                                        : umma_smem_desc_sw128(sA + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc_sw128(sB + k * args.mn_kstep, args.mn_lbo, args.mn_sbo)
                                        : umma_smem_desc_sw128(sB + k * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ======================================= epilogue ===========================================
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;   // row of the 128-row tile owned by this thread
    const int et = threadIdx.x - 64; // 0..127
    const bool store_thread = (et == 0);
    // bias exists only for forward Linears (B read K-major, segments tile N)
    const uint64_t dseed = drop_seed(args.drop);
    constexpr bool HAS_BIAS = !B_MN && (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_RELU);
    uint8_t* stg0 = smem + L::OFF_STG;
    uint8_t* stg1 = stg0 + STG_BYTES;
    int it = 0;
    uint32_t box_seq = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int n_blk = w % num_n;
      const int m_blk = (w / num_n) % num_m;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int n0 = n_blk * BN;
      const int m0 = m_blk * BM;
      const long long m = m0 + row;
      const bool row_ok = m < args.M;

      if (HAS_BIAS) {
        const int seg = n0 / args.b_seg_rows;
        const __nv_bfloat16* bp = args.bias[seg];
        for (int i = et; i < BN; i += 128) {
          const int n = n0 + i;
          s_bias[i] = (bp != nullptr && n < args.N) ? __bfloat162float(bp[n - seg * args.b_seg_rows]) : 0.f;
        }
        epi_bar_sync();
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

      if (EPI == EPI_REDUCE_F32) {
        // fp32 staging: 32 columns = 128 bytes per row; one TMA reduce-add box per 32 columns.
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          if (c == BN / 32 - 1) {
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
          }
          uint8_t* stg = (box_seq & 1u) ? stg1 : stg0;
          if (store_thread) tma_store_wait_read<1>();
          epi_bar_sync();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            stg_write16(stg, row, j, make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
          fence_proxy_async_smem();
          epi_bar_sync();
          if (store_thread) {
            tma_reduce_add_2d(&tm.d0, stg, n0 + c * 32, m0);
            tma_store_commit();
          }
          ++box_seq;
        }
      } else {
        // bf16 staging: 64 columns = 128 bytes per row; one TMA store box per 64 columns.
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          uint32_t o0[32];  // packed bf16 pairs: primary output, 64 columns
          uint32_t o1[(EPI == EPI_GELU) ? 32 : 1];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld32(t_row + c * 64 + h * 32, r);
            tmem_ld_wait();
            const int nb = n0 + c * 64 + h * 32;  // first global column of this 32-wide slab
            if (HAS_BIAS) {
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + s_bias[c * 64 + h * 32 + j]);
            }
            if (EPI == EPI_GELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float u0 = __uint_as_float(r[2 * j]), u1 = __uint_as_float(r[2 * j + 1]);
                o0[h * 16 + j] = pack_bf16x2(u0, u1);
                o1[h * 16 + j] = pack_bf16x2(gelu_erf(u0), gelu_erf(u1));
              }
            } else if (EPI == EPI_RELU) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t keep = 0xFFu;
                if (args.drop.p > 0.f)
                  keep = dropout_keep8(dseed, args.drop.site,
                                       (static_cast<uint64_t>(m) * args.N + nb + g * 8) >> 3, args.drop.thresh16);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float v0 = fmaxf(__uint_as_float(r[g * 8 + 2 * j]), 0.f);
                  float v1 = fmaxf(__uint_as_float(r[g * 8 + 2 * j + 1]), 0.f);
                  v0 = ((keep >> (2 * j)) & 1u) ? v0 * args.drop.scale : 0.f;
                  v1 = ((keep >> (2 * j + 1)) & 1u) ? v1 * args.drop.scale : 0.f;
                  o0[h * 16 + g * 4 + j] = pack_bf16x2(v0, v1);
                }
              }
            } else if (EPI == EPI_ADD || EPI == EPI_DGELU || EPI == EPI_DRELU) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 a = make_uint4(0, 0, 0, 0);
                if (row_ok && nb + g * 8 + 8 <= args.N)
                  a = __ldg(reinterpret_cast<const uint4*>(args.aux + m * args.ld_aux + nb + g * 8));
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 x = unpack_bf16x2(aw[j]);
                  float v0 = __uint_as_float(r[g * 8 + 2 * j]), v1 = __uint_as_float(r[g * 8 + 2 * j + 1]);
                  if (EPI == EPI_ADD) {
                    v0 += x.x;
                    v1 += x.y;
                  } else if (EPI == EPI_DGELU) {
                    v0 *= gelu_erf_grad(x.x);
                    v1 *= gelu_erf_grad(x.y);
                  } else {
                    v0 = x.x > 0.f ? v0 * args.relu_scale : 0.f;
                    v1 = x.y > 0.f ? v1 * args.relu_scale : 0.f;
                  }
                  o0[h * 16 + g * 4 + j] = pack_bf16x2(v0, v1);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                o0[h * 16 + j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
          }
          if (c == BN / 64 - 1) {
            // all TMEM reads of this accumulator are done: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
          }
          if (EPI == EPI_GELU) {
            if (store_thread) tma_store_wait_read<0>();
            epi_bar_sync();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              stg_write16(stg0, row, j, make_uint4(o0[4 * j], o0[4 * j + 1], o0[4 * j + 2], o0[4 * j + 3]));
              stg_write16(stg1, row, j, make_uint4(o1[4 * j], o1[4 * j + 1], o1[4 * j + 2], o1[4 * j + 3]));
            }
            fence_proxy_async_smem();
            epi_bar_sync();
            if (store_thread) {
              tma_store_2d(&tm.d0, stg0, n0 + c * 64, m0);
              tma_store_2d(&tm.d1, stg1, n0 + c * 64, m0);
              tma_store_commit();
            }
          } else {
            uint8_t* stg = (box_seq & 1u) ? stg1 : stg0;
            if (store_thread) tma_store_wait_read<1>();
            epi_bar_sync();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              stg_write16(stg, row, j, make_uint4(o0[4 * j], o0[4 * j + 1], o0[4 * j + 2], o0[4 * j + 3]));
            fence_proxy_async_smem();
            epi_bar_sync();
            if (store_thread) {
              tma_store_2d(&tm.d0, stg, n0 + c * 64, m0);
              tma_store_commit();
            }
            ++box_seq;
          }
        }
      }
    }
    if (store_thread) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_inst(const GemmTmaps& tm, const GemmArgs& args, int num_work, cudaStream_t stream) {
  constexpr int STAGES = (BN == 256) ? 4 : 6;
  using L = SmemLayout<BN, STAGES>;
  auto kfn = gemm_kernel<BN, A_MN, B_MN, EPI, STAGES>;
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    VLPK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
    attr_set = true;
  }
  const int grid = num_work < num_sms() ? num_work : num_sms();
  LaunchScope scope(A_MN ? CAT_GEMM_WGRAD : (B_MN ? CAT_GEMM_DGRAD : CAT_GEMM_FWD), 2.0 * args.M * args.N * args.K, stream);
  kfn<<<grid, NUM_THREADS, L::DYN_BYTES, stream>>>(tm, args);
  VLPK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN>
static int dispatch(const GemmDesc& g, const GemmTmaps& tm, const GemmArgs& args, int num_work, cudaStream_t s) {
  if (!g.a_mn && !g.b_mn) {
    switch (g.epi) {
      case EPI_STORE: return launch_inst<BN, false, false, EPI_STORE>(tm, args, num_work, s);
      case EPI_GELU: return launch_inst<BN, false, false, EPI_GELU>(tm, args, num_work, s);
      case EPI_RELU: return launch_inst<BN, false, false, EPI_RELU>(tm, args, num_work, s);
      default: break;
    }
  } else if (!g.a_mn && g.b_mn) {
    switch (g.epi) {
      case EPI_STORE: return launch_inst<BN, false, true, EPI_STORE>(tm, args, num_work, s);
      case EPI_ADD: return launch_inst<BN, false, true, EPI_ADD>(tm, args, num_work, s);
      case EPI_DGELU: return launch_inst<BN, false, true, EPI_DGELU>(tm, args, num_work, s);
      case EPI_DRELU: return launch_inst<BN, false, true, EPI_DRELU>(tm, args, num_work, s);
      default: break;
    }
  } else if (g.a_mn && g.b_mn) {
    if (g.epi == EPI_REDUCE_F32) return launch_inst<BN, true, true, EPI_REDUCE_F32>(tm, args, num_work, s);
  }
  set_error("gemm: unsupported (a_mn=%d, b_mn=%d, epi=%d) combination", (int)g.a_mn, (int)g.b_mn, g.epi);
  return -1;
}

int launch_gemm(const GemmDesc& g, cudaStream_t stream) {
  VLPK_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  VLPK_CHECK_ARG(g.N % 8 == 0, "gemm: N=%d must be a multiple of 8", g.N);
  VLPK_CHECK_ARG(g.nseg >= 1 && g.nseg <= 3, "gemm: nseg=%d", g.nseg);
  int bn = g.bn;
  if (bn == 0) {
    // Heuristic: prefer the 256-wide tile (best smem-bandwidth : MMA ratio) unless it leaves the last
    // wave mostly idle; the hot-path shapes are tabulated in DESIGN.md.
    const int num_m = (g.M + BM - 1) / BM;
    const int t256 = num_m * ((g.N + 255) / 256) * g.splits;
    const int t128 = num_m * ((g.N + 127) / 128) * g.splits;
    const int sms = num_sms();
    const double e256 = double(t256) / (double((t256 + sms - 1) / sms) * sms);
    const double e128 = double(t128) / (double((t128 + sms - 1) / sms) * sms) * 0.92;  // 128-wide tile is ~8% less efficient
    bn = (g.N % 256 == 0 && e256 >= e128) ? 256 : 128;
    if (g.N % 128 != 0 && g.N % 256 != 0) bn = 128;
  }
  VLPK_CHECK_ARG(bn == 128 || bn == 256, "gemm: tile N %d unsupported", bn);

  const int seg_rows = g.nseg > 1 ? g.b_seg_rows : (g.b_mn ? g.K : g.N);
  if (g.nseg > 1) {
    VLPK_CHECK_ARG(seg_rows % (g.b_mn ? BK : bn) == 0, "gemm: segment rows %d not tile aligned", seg_rows);
  }
  const int total_kb = (g.K + BK - 1) / BK;
  int splits = g.splits < 1 ? 1 : g.splits;
  if (splits > total_kb) splits = total_kb;
  {
    const int kb_per = (total_kb + splits - 1) / splits;
    splits = (total_kb + kb_per - 1) / kb_per;  // no empty splits
  }
  VLPK_CHECK_ARG(splits == 1 || g.epi == EPI_REDUCE_F32, "gemm: split-K needs EPI_REDUCE_F32");

  GemmTmaps tm;
  memset(&tm, 0, sizeof(tm));
  if (!g.a_mn) {
    VLPK_TRY(make_tmap_2d(&tm.a, TM_BF16, g.A, g.K, g.M, g.lda, BK, BM));
  } else {
    VLPK_TRY(make_tmap_2d(&tm.a, TM_BF16, g.A, g.M, g.K, g.lda, 64, BK));
  }
  for (int s = 0; s < g.nseg; ++s) {
    if (!g.b_mn) {
      const int rows = g.nseg > 1 ? seg_rows : g.N;
      VLPK_TRY(make_tmap_2d(&tm.b[s], TM_BF16, g.B[s], g.K, rows, g.ldb, BK, bn));
    } else {
      const int rows = g.nseg > 1 ? seg_rows : g.K;
      VLPK_TRY(make_tmap_2d(&tm.b[s], TM_BF16, g.B[s], g.N, rows, g.ldb, 64, BK));
    }
  }
  for (int s = g.nseg; s < 3; ++s) tm.b[s] = tm.b[0];
  if (g.epi == EPI_REDUCE_F32) {
    VLPK_TRY(make_tmap_2d(&tm.d0, TM_F32, g.D0, g.N, g.M, g.ldd0, 32, BM));
    tm.d1 = tm.d0;
  } else {
    VLPK_TRY(make_tmap_2d(&tm.d0, TM_BF16, g.D0, g.N, g.M, g.ldd0, 64, BM));
    if (g.epi == EPI_GELU) {
      VLPK_CHECK_ARG(g.D1 != nullptr, "gemm: EPI_GELU needs D1");
      VLPK_TRY(make_tmap_2d(&tm.d1, TM_BF16, g.D1, g.N, g.M, g.ldd1, 64, BM));
    } else {
      tm.d1 = tm.d0;
    }
  }
  if (g.epi == EPI_ADD || g.epi == EPI_DGELU || g.epi == EPI_DRELU) {
    VLPK_CHECK_ARG(g.aux != nullptr && (g.ld_aux % 8) == 0 && (reinterpret_cast<uintptr_t>(g.aux) & 15u) == 0,
                   "gemm: aux must be 16-byte aligned with ld %% 8 == 0");
  }

  GemmArgs a;
  a.M = g.M;
  a.N = g.N;
  a.K = g.K;
  a.b_seg_rows = seg_rows;
  a.splits = splits;
  for (int s = 0; s < 3; ++s) a.bias[s] = (s < g.nseg) ? g.bias[s] : nullptr;
  a.aux = g.aux;
  a.ld_aux = g.ld_aux;
  a.relu_scale = g.relu_scale;
  a.drop = g.drop;
  a.mn_lbo = g_mn_lbo;
  a.mn_sbo = g_mn_sbo;
  a.mn_kstep = g_mn_kstep;

  const int num_m = (g.M + BM - 1) / BM;
  const int num_n = (g.N + bn - 1) / bn;
  const int num_work = num_m * num_n * splits;
  if (bn == 256) return dispatch<256>(g, tm, a, num_work, stream);
  return dispatch<128>(g, tm, a, num_work, stream);
}

}  // namespace vlpk
