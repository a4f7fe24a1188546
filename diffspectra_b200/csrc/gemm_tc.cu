// Hand-written sm_100a GEMM:  out[M,N] = act(A[M,K] * W[N,K]^T + bias + addmat),  bf16 in, fp32 accumulate.
//
// Persistent, warp-specialised:  warp 4 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B tiles, mbarrier
// pipeline),  warp 5 = tcgen05.mma issuer + TMEM owner (one elected thread, accumulators double-buffered in
// TMEM so the epilogue of tile i overlaps the MMAs of tile i+1),  warps 0-3 = epilogue (tcgen05.ld -> bias /
// activation -> bf16|fp32 global stores, one output row per thread).
//
// This one kernel serves every dense contraction of DMT / SpecFormer in bf16 mode (reference: all nn.Linear
// calls of models/dmt.py, models/layers.py, models/specformer.py — SURVEY.md §2.1 'addmm' row).
#include "context.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int BM = 128;   // UMMA M (cta_group::1)
constexpr int BK = 64;    // one 128-byte swizzle atom of bf16
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 4;
constexpr int kThreads = (kEpiWarps + 2) * 32;

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kWBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct Epi {
  const float* bias;
  const float* addmat;
  void* out;
  int ldo, ldadd;
  int out_dtype;
  int act;
};

template <int CH>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[CH], const Epi& ep, int row, int col0, int N) {
  float f[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) f[i] = __uint_as_float(v[i]);
  if (ep.bias) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (col0 + i < N) f[i] += __ldg(ep.bias + col0 + i);
  }
  if (ep.addmat) {
    const float* ar = ep.addmat + static_cast<size_t>(row) * ep.ldadd + col0;
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (col0 + i < N) f[i] += __ldg(ar + i);
  }
  if (ep.act != ACT_NONE) {
#pragma unroll
    for (int i = 0; i < CH; ++i) f[i] = apply_act<true>(f[i], ep.act);
  }
  const bool full = (col0 + CH <= N);
  if (ep.out_dtype == DT_BF16) {
    bf16* o = reinterpret_cast<bf16*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < CH; i += 8) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(f[i], f[i + 1]);
        __nv_bfloat162 p1 = __floats2bfloat162_rn(f[i + 2], f[i + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(f[i + 4], f[i + 5]);
        __nv_bfloat162 p3 = __floats2bfloat162_rn(f[i + 6], f[i + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        u.z = *reinterpret_cast<uint32_t*>(&p2);
        u.w = *reinterpret_cast<uint32_t*>(&p3);
        *reinterpret_cast<uint4*>(o + i) = u;
      }
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (col0 + i < N) o[i] = __float2bfloat16_rn(f[i]);
    }
  } else {
    float* o = reinterpret_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < CH; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (col0 + i < N) o[i] = f[i];
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, Epi ep, int M, int N,
               int K) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smW = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + C::kStages;         // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * C::kStages;     // [2]        MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * C::kStages + 2;   // [2]     epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + BK - 1) / BK;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int i = 0; i < C::kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], kEpiWarps * 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 5) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BM;
        const int n0 = (t % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
          ptx::tma_load_2d(smA + stage * C::kABytes, &tmA, &full_bar[stage], kb * BK, m0);
          ptx::tma_load_2d(smW + stage * C::kWBytes, &tmW, &full_bar[stage], kb * BK, n0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smA + stage * C::kABytes);
          const uint32_t w_addr = ptx::smem_u32(smW + stage * C::kWBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = ptx::umma_smem_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t wd = ptx::umma_smem_desc_sw128(w_addr + k * UMMA_K * 2);
            ptx::umma_bf16(d_tmem, ad, wd, idesc, (kb | k) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);                  // smem slot free once these MMAs retire
          if (kb == k_blocks - 1) ptx::umma_commit(&tfull_bar[as]);   // accumulator ready
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 0..3 <-> TMEM lane quadrants 0..3) =====================
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / n_tiles) * BM;
      const int n0 = (t % n_tiles) * BN;
      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();
      const int row = m0 + warp * 32 + lane;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(as * BN);
      if constexpr (BN >= 32) {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          ptx::tmem_ld32(t_addr + c, v);
          ptx::tmem_ld_wait();
          if (row < M && n0 + c < N) epilogue_chunk<32>(v, ep, row, n0 + c, N);
        }
      } else {
        uint32_t v[16];
        ptx::tmem_ld16(t_addr, v);
        ptx::tmem_ld_wait();
        if (row < M) epilogue_chunk<16>(v, ep, row, n0, N);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(DsContext* ctx, CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DS_CHECK(r == CUDA_SUCCESS, DS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d base=%p", (int)r,
           rows, cols, ld, base);
  return DS_OK;
}

template <int BN>
int launch_bn(DsContext* ctx, const GemmDesc& g, cudaStream_t s) {
  using C = Cfg<BN>;
  CUtensorMap tmA, tmW;
  DS_TRY(make_tmap(ctx, &tmA, g.A, g.M, g.K, g.lda, BM));
  DS_TRY(make_tmap(ctx, &tmW, g.W, g.N, g.K, g.ldw, BN));
  Epi ep{g.bias, g.addmat, g.out, g.ldo, g.ldadd, g.out_dtype, g.act};
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  gemm_tc_kernel<BN><<<grid, kThreads, C::kSmemBytes, s>>>(tmA, tmW, ep, g.M, g.N, g.K);
  DS_CUDA_CHECK(cudaGetLastError());
  ctx->launch_count++;
  return DS_OK;
}

template <int BN>
int set_attr() {
  DS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes));
  return DS_OK;
}

}  // namespace

int gemm_tc_init(DsContext* ctx) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DS_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  DS_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, DS_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  ctx->encode_tiled = fn;
  DS_TRY(set_attr<16>());
  DS_TRY(set_attr<32>());
  DS_TRY(set_attr<64>());
  DS_TRY(set_attr<128>());
  DS_TRY(set_attr<256>());
  return DS_OK;
}

int gemm_tc_launch(DsContext* ctx, const GemmDesc& g, cudaStream_t s) {
  DS_CHECK(g.a_dtype == DT_BF16, DS_ERR_INVALID, "gemm_tc: A/W must be bf16");
  DS_CHECK(g.M > 0 && g.N > 0 && g.K > 0, DS_ERR_INVALID, "gemm_tc: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  DS_CHECK((g.lda % 8) == 0 && (g.ldw % 8) == 0, DS_ERR_INVALID, "gemm_tc: lda/ldw must be multiples of 8 (16 B rows)");
  DS_CHECK((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.W) & 15) == 0, DS_ERR_INVALID,
           "gemm_tc: A/W must be 16-byte aligned");
  if (g.N <= 16) return launch_bn<16>(ctx, g, s);
  if (g.N <= 32) return launch_bn<32>(ctx, g, s);
  if (g.N <= 64) return launch_bn<64>(ctx, g, s);
  if (g.N <= 128) return launch_bn<128>(ctx, g, s);
  return launch_bn<256>(ctx, g, s);
}
