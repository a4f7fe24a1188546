// Hand-written sm_100a GEMM:  out[M,N] = epilogue(A[M,K] * W[N,K]^T),  bf16 in, fp32 accumulate in TMEM.
//
// Persistent, warp-specialised, 320 threads:
//   warp 8      TMA producer  (cp.async.bulk.tensor, SWIZZLE_128B tiles, mbarrier ring of kStages slots)
//   warp 9      tcgen05.mma issuer + TMEM owner (one elected thread; TWO accumulator stages in TMEM)
//   warps 0-3   epilogue group 0  (accumulator stage 0, even tiles)
//   warps 4-7   epilogue group 1  (accumulator stage 1, odd tiles)
// Each epilogue thread owns one full output row of its tile (tcgen05.ld 32x32b), so row-wise epilogues
// (LayerNorm, dot products) need no cross-thread traffic.  bf16 outputs are staged in swizzled shared memory
// and written with TMA stores (coalesced, bounds-clipped); the two groups let the epilogue of tile i overlap
// the MMAs of tiles i+1, i+2.
//
// Epilogue modes (fusions of the reference's elementwise ops into the producing contraction):
//   STORE    act(acc + bias + addmat)                                  every nn.Linear (+SiLU/tanh/GELU)
//   LNMOD    modulate(LayerNorm(acc + bias), shift[mol], scale[mol])    dmt.py:139,149 (edge_emb -> norm1_edge)
//   RESGATE  resid + gate[mol] * (acc + bias) -> fp32 stream + bf16 copy  dmt.py:162-163,168-169 (FFN residuals)
//   COORD    mean(tanh(W2 . SiLU(2 (acc + bias))) * [1, adj2d, adjsp])   dmt.py:32-35,45-51 (coord_mlp + heads; W, bias pre-halved)
//   EHEAD    [w_e . SiLU(acc[:32] + b), w_t . SiLU(acc[32:] + b)]          dmt.py:234-247,394 (edge heads, layers 2+4)
#include "context.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int BM = 128;   // UMMA M (cta_group::1)
constexpr int BK = 64;    // one 128-byte swizzle atom of bf16
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
// COORD mode is bound by its epilogue (7 CUDA-core instructions per accumulator element on two warps per scheduler), so it
// gets eight more epilogue warps: warps 10-17 take columns [128, 256) of the rows that warps 0-7 share with them
template <int MODE>
constexpr int threads_for() { return (kEpiWarps + 2 + (MODE == GEMM_COORD ? 8 : 0)) * 32; }
constexpr int kStageBox = 32 * 128;   // one staging box: 32 rows x 128 bytes

template <int BN, int MODE, bool TMA_OUT>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kWBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kWBytes;
  // per-warp staging: bf16 TMA-store boxes (2 x 4 KB), RESGATE: fp32 in/out tile (2 boxes) + one bf16 box
  static constexpr int kWarpStaging = (MODE == GEMM_RESGATE) ? 3 * kStageBox : (TMA_OUT ? 2 * kStageBox : 0);
  static constexpr int kStagingBytes = kEpiWarps * kWarpStaging;
  static constexpr int kAuxBytes = 2 * 256 * 4 /*bias, per group*/ + (MODE == GEMM_COORD ? 2 * 256 * 16 : 0) + 256 /*barriers*/;
  static constexpr int kBudget = 220 * 1024 - kStagingBytes - kAuxBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kAuxBytes + 1024 /*align slack*/;
};

struct Epi {
  const float* bias;
  const float* addmat;
  void* out;            // direct-store output (fp32, or bf16 when not TMA_OUT)
  int ldo, ldadd;
  int out_dtype;
  int act;
  // row -> molecule -> adaLN row
  const uint32_t* row_info;
  int info_shift;
  const float* ada;     // pre-offset to the block's base; row stride ADA_LD
  int off_a, off_b;     // LNMOD: shift, scale offsets; RESGATE: off_a = gate offset
  const float* resid;
  int ldres;
  const float* wc2;     // COORD: [3,256]
  const uint8_t* pflags;
  float* wdir;
  int out2_present;
  int split_n;          // STORE + TMA_OUT: column tiles at n0 >= split_n are stored through the second map (0 = off)
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// write 64 bf16 (one 128-byte row) of staging row `r` (0..31) with the SWIZZLE_128B pattern
__device__ __forceinline__ void stage_row_bf16(uint8_t* box, int r, const float (&f)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 u;
    u.x = pack_bf16(f[c * 8 + 0], f[c * 8 + 1]);
    u.y = pack_bf16(f[c * 8 + 2], f[c * 8 + 3]);
    u.z = pack_bf16(f[c * 8 + 4], f[c * 8 + 5]);
    u.w = pack_bf16(f[c * 8 + 6], f[c * 8 + 7]);
    *reinterpret_cast<uint4*>(box + r * 128 + ((c ^ (r & 7)) << 4)) = u;
  }
}

// 32 fp32 (one 128-byte row) of staging row r, SWIZZLE_128B pattern (chunk = 4 floats)
__device__ __forceinline__ void stage_row_f32(uint8_t* box, int r, const float* f) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(box + r * 128 + ((c ^ (r & 7)) << 4)) = make_float4(f[c * 4], f[c * 4 + 1], f[c * 4 + 2], f[c * 4 + 3]);
}
__device__ __forceinline__ void unstage_row_f32(const uint8_t* box, int r, float* f) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(box + r * 128 + ((c ^ (r & 7)) << 4));
    f[c * 4] = v.x; f[c * 4 + 1] = v.y; f[c * 4 + 2] = v.z; f[c * 4 + 3] = v.w;
  }
}

// direct (register -> global) store of CH consecutive columns of one row
template <int CH>
__device__ __forceinline__ void direct_store(const float (&f)[CH], void* out, int out_dtype, int ldo, int row, int col0, int N) {
  const bool full = (col0 + CH <= N);
  if (out_dtype == DT_BF16) {
    bf16* o = reinterpret_cast<bf16*>(out) + static_cast<size_t>(row) * ldo + col0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < CH; i += 8) {
        uint4 u;
        u.x = pack_bf16(f[i], f[i + 1]); u.y = pack_bf16(f[i + 2], f[i + 3]);
        u.z = pack_bf16(f[i + 4], f[i + 5]); u.w = pack_bf16(f[i + 6], f[i + 7]);
        *reinterpret_cast<uint4*>(o + i) = u;
      }
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (col0 + i < N) o[i] = __float2bfloat16_rn(f[i]);
    }
  } else {
    float* o = reinterpret_cast<float*>(out) + static_cast<size_t>(row) * ldo + col0;
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

template <int CH>
__device__ __forceinline__ void load_acc(uint32_t taddr, float (&f)[CH]) {
  uint32_t v[CH];
  if constexpr (CH == 64) ptx::tmem_ld64_sync(taddr, v);
  else if constexpr (CH == 32) ptx::tmem_ld32_sync(taddr, v);
  else ptx::tmem_ld16_sync(taddr, v);
#pragma unroll
  for (int i = 0; i < CH; ++i) f[i] = __uint_as_float(v[i]);
}

template <int BN, int MODE, bool TMA_OUT>
__global__ void __launch_bounds__(threads_for<MODE>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmF, Epi ep, int M, int N, int K) {
  using C = Cfg<BN, MODE, TMA_OUT>;
  // 1024-byte alignment is requested from the compiler/driver (no static shared memory in this kernel), so plain
  // pointer arithmetic keeps the shared address space visible and the epilogues compile to LDS/STS, not generic LD/ST.
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  uint8_t* smA = smem;
  uint8_t* smW = smem + C::kStages * C::kABytes;
  uint8_t* staging = smem + C::kStages * C::kStageBytes;                 // 1024-aligned (stage sizes are multiples of 1 KB)
  float* sbias = reinterpret_cast<float*>(staging + C::kStagingBytes);   // [2][256]
  float4* swc2 = reinterpret_cast<float4*>(sbias + 512);                 // COORD: [256] (w0, w1, w2, bias)
  float4* spart = swc2 + 256;                                            // COORD: [2][128] partial sums of columns 128..255
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sbias) + 2048 + (MODE == GEMM_COORD ? 8192 : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tfull_bar = bars + 2 * C::kStages;
  uint64_t* tempty_bar = bars + 2 * C::kStages + 2;
  uint64_t* rbar = bars + 2 * C::kStages + 4;         // [8] RESGATE: residual tile landed (one per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + BK - 1) / BK;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 8 && lane == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();      // SWIZZLE_128B tiles need a 1024-byte aligned base
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    if (TMA_OUT) ptx::prefetch_tmap(&tmO);
    for (int i = 0; i < C::kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], MODE == GEMM_COORD ? 256 : 128);
    }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&rbar[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  if constexpr (MODE == GEMM_COORD) {
    if (threadIdx.x < 256) {     // per column PAIR (2p, 2p+1): [2p] = (w0, w0', w1, w1'), [2p+1] = (w2, w2', bias, bias')
      const int c2 = threadIdx.x & ~1;
      swc2[threadIdx.x] = (threadIdx.x & 1) ? make_float4(ep.wc2[512 + c2], ep.wc2[513 + c2], ep.bias[c2], ep.bias[c2 + 1])
                                            : make_float4(ep.wc2[c2], ep.wc2[c2 + 1], ep.wc2[256 + c2], ep.wc2[257 + c2]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above (barriers, TMEM, weight-only reads) overlapped the predecessor's tail

  if (warp == 8) {
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
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
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
          ptx::umma_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) ptx::umma_commit(&tfull_bar[as]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: group g owns accumulator stage g and every second tile =====================
    const bool second = MODE == GEMM_COORD && warp >= 10;          // COORD: column half [128, 256)
    const int g = MODE == GEMM_COORD ? (((warp < 8 ? warp : warp - 10) >> 2) & 1) : (warp >> 2), wq = warp & 3;
    const int gtid = threadIdx.x & 127;
    float* gb = sbias + g * 256;
    uint8_t* my_stage = staging + warp * C::kWarpStaging;
    uint32_t rphase = 0;
    int sb = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      if ((it & 1) != g) continue;
      const uint32_t gphase = (it >> 1) & 1;
      const int m0 = (t / n_tiles) * BM;
      const int n0 = (t % n_tiles) * BN;
      if constexpr (MODE != GEMM_COORD) {
        named_bar_sync(1 + g, 128);                  // everyone is done with the previous tile's bias
        for (int i = gtid; i < BN; i += 128) gb[i] = (ep.bias && n0 + i < N) ? ep.bias[n0 + i] : 0.f;
        named_bar_sync(1 + g, 128);
      }
      // the row's molecule and its adaLN vectors are known before the accumulator is: fetch the index and pull the
      // lines into L1 now, so that the epilogue does not start with two dependent L2 round trips
      uint32_t mol_pre = 0u;
      if constexpr (MODE == GEMM_LNMOD || MODE == GEMM_RESGATE) {
        const int row_pre = m0 + wq * 32 + lane;
        if (row_pre < M && ep.row_info != nullptr && ep.ada != nullptr) {
          mol_pre = ep.row_info[row_pre] >> ep.info_shift;
          const float* ar = ep.ada + static_cast<size_t>(mol_pre) * ADA_LD;
          constexpr int kLines = MODE == GEMM_LNMOD ? 2 : BN / 32;      // 128-byte lines of the first vector
#pragma unroll
          for (int i = 0; i < kLines; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(ar + ep.off_a + i * 32));
          if constexpr (MODE == GEMM_LNMOD) {
#pragma unroll
            for (int i = 0; i < 2; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(ar + ep.off_b + i * 32));
          }
        }
      }
      uint8_t fl_pre = 0;
      if constexpr (MODE == GEMM_COORD) {              // adjacency bits of the row: known now, needed at the very end
        const int row_pre = m0 + wq * 32 + lane;
        if (row_pre < M && warp < 8) fl_pre = ep.pflags[row_pre];
      }
      ptx::mbar_wait(&tfull_bar[g], gphase);
      ptx::tc_fence_after();
      const int row = m0 + wq * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(g * BN);

      if constexpr (MODE == GEMM_STORE) {
        constexpr int CH = BN >= 64 ? (TMA_OUT ? 64 : 32) : BN;
#pragma unroll 1
        for (int c = 0; c < BN; c += CH) {
          float f[CH];
          load_acc<CH>(t_addr + c, f);
          if (ep.bias) {                            // this epilogue is bound by issue slots (~3 us per instruction per element at
#pragma unroll                                      // [162 k, 512]): no add for a Linear without bias, packed adds otherwise
            for (int i = 0; i + 1 < CH; i += 2) {
              const float2 v = fadd2(make_float2(f[i], f[i + 1]), *reinterpret_cast<const float2*>(gb + c + i));
              f[i] = v.x;
              f[i + 1] = v.y;
            }
          }
          if (ep.addmat && row_ok) {
            const float* ar = ep.addmat + static_cast<size_t>(row) * ep.ldadd + n0 + c;
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (n0 + c + i < N) f[i] += __ldg(ar + i);
          }
          if (ep.act == ACT_TANH_MIX) {             // MUFU.TANH-bound epilogue: every second column pair on the FMA pipe
#pragma unroll
            for (int i = 0; i + 3 < CH; i += 4) {
              const float2 pv = tanh_poly2(make_float2(f[i], f[i + 1]));
              f[i] = pv.x;
              f[i + 1] = pv.y;
              f[i + 2] = act_tanh<true>(f[i + 2]);
              f[i + 3] = act_tanh<true>(f[i + 3]);
            }
          } else if (ep.act != ACT_NONE) {
#pragma unroll
            for (int i = 0; i < CH; ++i) f[i] = apply_act<true>(f[i], ep.act);
          }
          if constexpr (TMA_OUT) {
            if (lane == 0) tma_store_wait_read<1>();       // the box we are about to overwrite has been read
            __syncwarp();
            uint8_t* box = my_stage + sb * kStageBox;
            stage_row_bf16(box, lane, f);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (n0 + c < N && m0 + wq * 32 < M) {
                if (ep.split_n > 0 && n0 >= ep.split_n) tma_store_2d(&tmR, box, n0 + c - ep.split_n, m0 + wq * 32);   // second destination
                else tma_store_2d(&tmO, box, n0 + c, m0 + wq * 32);
              }
              tma_store_commit();
            }
            sb ^= 1;
          } else {
            if (row_ok && n0 + c < N) direct_store<CH>(f, ep.out, ep.out_dtype, ep.ldo, row, n0 + c, N);
          }
        }
      } else if constexpr (MODE == GEMM_LNMOD) {
        // BN == N == 64: the thread holds the whole row
        float f[64];
        load_acc<64>(t_addr, f);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) { f[i] += gb[i]; s += f[i]; }
        const float mean = s * (1.0f / 64.0f);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) { f[i] -= mean; q += f[i] * f[i]; }
        const float is = rsqrtf(q * (1.0f / 64.0f) + 1e-6f);
        const uint32_t mol = mol_pre;
        const float* ar = ep.ada + static_cast<size_t>(mol) * ADA_LD;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float4 sh = __ldg(reinterpret_cast<const float4*>(ar + ep.off_a + i));
          const float4 sc = __ldg(reinterpret_cast<const float4*>(ar + ep.off_b + i));
          f[i + 0] = (f[i + 0] * is) * (1.0f + sc.x) + sh.x;
          f[i + 1] = (f[i + 1] * is) * (1.0f + sc.y) + sh.y;
          f[i + 2] = (f[i + 2] * is) * (1.0f + sc.z) + sh.z;
          f[i + 3] = (f[i + 3] * is) * (1.0f + sc.w) + sh.w;
        }
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        uint8_t* box = my_stage + sb * kStageBox;
        stage_row_bf16(box, lane, f);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (m0 + wq * 32 < M) tma_store_2d(&tmO, box, 0, m0 + wq * 32);
          tma_store_commit();
        }
        sb ^= 1;
      } else if constexpr (MODE == GEMM_RESGATE) {
        // out = resid + gate[mol] * (acc + bias): residual tile in via TMA (issued before the accumulator wait),
        // fp32 stream and bf16 copy out via TMA stores; resid == null (tmR unused): plain acc + bias.
        const bool active = m0 + wq * 32 < M;
        const bool has_res = ep.resid != nullptr;
        const uint32_t mol = (row_ok && has_res) ? mol_pre : 0u;
        const float* gr = has_res ? ep.ada + static_cast<size_t>(mol) * ADA_LD + ep.off_a : nullptr;
        uint8_t* fbox = my_stage;                    // two fp32 boxes (cols 0-31, 32-63 of the chunk)
        uint8_t* bbox = my_stage + 2 * kStageBox;    // one bf16 box
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
          if (active && has_res) {
            if (lane == 0) {
              tma_store_wait_read<0>();              // previous stores have finished reading the staging boxes
              ptx::mbar_arrive_expect_tx(&rbar[warp], 2 * kStageBox);
              ptx::tma_load_2d(fbox, &tmR, &rbar[warp], n0 + c, m0 + wq * 32);
              ptx::tma_load_2d(fbox + kStageBox, &tmR, &rbar[warp], n0 + c + 32, m0 + wq * 32);
            }
          } else if (lane == 0) {
            tma_store_wait_read<0>();
          }
          __syncwarp();
          float f[64];
          load_acc<64>(t_addr + c, f);
          if (active) {
            if (has_res) {
              ptx::mbar_wait(&rbar[warp], rphase);
              rphase ^= 1;
              float r[64];
              unstage_row_f32(fbox, lane, r);
              unstage_row_f32(fbox + kStageBox, lane, r + 32);
#pragma unroll
              for (int i = 0; i < 64; i += 4) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(gr + n0 + c + i));
                f[i + 0] = r[i + 0] + g4.x * (f[i + 0] + gb[c + i + 0]);
                f[i + 1] = r[i + 1] + g4.y * (f[i + 1] + gb[c + i + 1]);
                f[i + 2] = r[i + 2] + g4.z * (f[i + 2] + gb[c + i + 2]);
                f[i + 3] = r[i + 3] + g4.w * (f[i + 3] + gb[c + i + 3]);
              }
              __syncwarp();                          // every lane has read its residual row
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i) f[i] += gb[c + i];
            }
            stage_row_f32(fbox, lane, f);
            stage_row_f32(fbox + kStageBox, lane, f + 32);
            stage_row_bf16(bbox, lane, f);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmF, fbox, n0 + c, m0 + wq * 32);
              tma_store_2d(&tmF, fbox + kStageBox, n0 + c + 32, m0 + wq * 32);
              if (ep.out2_present) tma_store_2d(&tmO, bbox, n0 + c, m0 + wq * 32);
              tma_store_commit();
            }
          }
        }
      } else if constexpr (MODE == GEMM_EHEAD) {
        // BN == N == 64: second layers of edge_exist_mlp | edge_type_mlp (block-diagonal weight), SiLU, then the
        // two 32 -> 1 output layers as per-row dot products  (dmt.py:234-247,394)
        float f[64];
        load_acc<64>(t_addr, f);
        float s0 = ep.wc2[64], s1 = ep.wc2[65];       // output biases
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s0 = fmaf(act_silu<true>(f[i] + gb[i]), __ldg(ep.wc2 + i), s0);
          s1 = fmaf(act_silu<true>(f[32 + i] + gb[32 + i]), __ldg(ep.wc2 + 32 + i), s1);
        }
        if (row_ok) *reinterpret_cast<float2*>(ep.wdir + static_cast<size_t>(row) * 2) = make_float2(s0, s1);
      } else {   // GEMM_COORD, BN == N == 256
        // two columns per step as packed fp32 (FADD2 + 4 FFMA2 + 2 MUFU per pair instead of 2 FADD + 8 FFMA + 2 MUFU):
        // even / odd columns accumulate separately and meet at the end
        float2 p0 = make_float2(0.f, 0.f), p1 = p0, p2 = p0;
        const int c_lo = second ? 128 : 0;
#pragma unroll 1
        for (int c = c_lo; c < c_lo + 128; c += 32) {
          float f[32];
          load_acc<32>(t_addr + c, f);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float4 wa = swc2[c + i], wb = swc2[c + i + 1];
            const float2 h = fadd2(make_float2(f[i], f[i + 1]), make_float2(wb.z, wb.w));   // coord_mlp.0 is packed with 0.5 W, 0.5 b
            const float2 v = ffma2(h, make_float2(act_tanh<true>(h.x), act_tanh<true>(h.y)), h);   // SiLU(2h) = h + h tanh(h)
            p0 = ffma2(v, make_float2(wa.x, wa.y), p0);
            p1 = ffma2(v, make_float2(wa.z, wa.w), p1);
            p2 = ffma2(v, make_float2(wb.x, wb.y), p2);
          }
        }
        float s0 = p0.x + p0.y, s1 = p1.x + p1.y, s2 = p2.x + p2.y;
        float4* sp = spart + g * 128 + wq * 32 + lane;
        if (second) *sp = make_float4(s0, s1, s2, 0.f);
        named_bar_sync(3 + g, 256);                     // the two warps of every row have met
        if (!second) {
          const float4 o = *sp;
          s0 += o.x; s1 += o.y; s2 += o.z;
        }
        if (row_ok && !second) {
          const uint8_t fl = fl_pre;                    // per directed edge (k_dir_flags), fetched before the accumulator wait
          const float a2 = (fl & 1) ? 1.f : 0.f, asp = (fl & 2) ? 1.f : 0.f;
          ep.wdir[row] = (act_tanh<true>(s0) + act_tanh<true>(s1) * a2 + act_tanh<true>(s2) * asp) / 3.0f;
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty_bar[g]);
    }
    if ((TMA_OUT || MODE == GEMM_RESGATE) && lane == 0) tma_store_wait_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(DsContext* ctx, CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_cols, int box_rows,
              bool f32 = false) {
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  const int es = f32 ? 4 : 2;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                  gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DS_CHECK(r == CUDA_SUCCESS, DS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d base=%p", (int)r,
           rows, cols, ld, base);
  return DS_OK;
}

template <int BN, int MODE, bool TMA_OUT>
int launch_cfg(DsContext* ctx, const GemmDesc& g, cudaStream_t s) {
  using C = Cfg<BN, MODE, TMA_OUT>;
  static bool attr_set[64] = {};            // the attribute is per device: one flag per device ordinal
  if (!attr_set[ctx->device & 63]) {
    DS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE, TMA_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set[ctx->device & 63] = true;
  }
  CUtensorMap tmA, tmW, tmO, tmR, tmF;
  DS_TRY(make_tmap(ctx, &tmA, g.A, g.M, g.K, g.lda, BK, BM));
  DS_TRY(make_tmap(ctx, &tmW, g.W, g.N, g.K, g.ldw, BK, BN));
  tmO = tmA; tmR = tmA; tmF = tmA;
  if (MODE == GEMM_RESGATE) {
    DS_TRY(make_tmap(ctx, &tmF, g.out, g.M, g.N, g.ldo, 32, 32, true));
    if (g.resid) DS_TRY(make_tmap(ctx, &tmR, g.resid, g.M, g.N, g.ldres, 32, 32, true));
    if (g.out2) DS_TRY(make_tmap(ctx, &tmO, g.out2, g.M, g.N, g.ldo2, 64, 32));
  } else if (TMA_OUT) {
    if (g.split_n > 0) {      // two destinations: columns [0, split_n) -> out, [split_n, N) -> out2
      DS_CHECK(MODE == GEMM_STORE && g.out2 != nullptr && (g.split_n % BN) == 0 && ((g.N - g.split_n) % 8) == 0 && (g.ldo2 % 8) == 0 &&
                   (reinterpret_cast<uintptr_t>(g.out2) & 15) == 0,
               DS_ERR_INVALID, "gemm_tc: split output needs split_n on a column-tile boundary and a 16-byte aligned second destination");
      DS_TRY(make_tmap(ctx, &tmO, g.out, g.M, g.split_n, g.ldo, 64, 32));
      DS_TRY(make_tmap(ctx, &tmR, g.out2, g.M, g.N - g.split_n, g.ldo2, 64, 32));
    } else {
      DS_TRY(make_tmap(ctx, &tmO, g.out, g.M, g.N, g.ldo, 64, 32));
    }
  }
  Epi ep;
  ep.bias = g.bias; ep.addmat = g.addmat; ep.out = g.out; ep.ldo = g.ldo; ep.ldadd = g.ldadd;
  ep.out_dtype = g.out_dtype; ep.act = g.act;
  ep.row_info = g.row_info; ep.info_shift = g.info_shift; ep.ada = g.ada; ep.off_a = g.off_a; ep.off_b = g.off_b;
  ep.resid = g.resid; ep.ldres = g.ldres; ep.wc2 = g.wc2; ep.pflags = g.pflags; ep.wdir = g.wdir;
  ep.out2_present = g.out2 != nullptr;
  ep.split_n = (TMA_OUT && MODE == GEMM_STORE) ? g.split_n : 0;
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int sms = (ctx->cta_cap > 0 && ctx->cta_cap < ctx->num_sms) ? ctx->cta_cap : ctx->num_sms;
  const int grid = tiles < sms ? tiles : sms;
  if (g_ds_prof.on)   // shape tag for ds_profile_end: mode(3) | N(15) | K(16) | M(30)
    g_ds_prof.tag = (static_cast<long long>(MODE) << 61) | (static_cast<long long>(g.N) << 46) | (static_cast<long long>(g.K) << 30) | g.M;
  ds_launch(gemm_tc_kernel<BN, MODE, TMA_OUT>, dim3(grid), dim3(threads_for<MODE>()), C::kSmemBytes, s, tmA, tmW, tmO, tmR, tmF, ep, g.M, g.N, g.K);
  DS_CUDA_CHECK(cudaGetLastError());
  ctx->launch_count++;
  return DS_OK;
}

}  // namespace

int ds_make_tmap_2d(DsContext* ctx, CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_cols,
                    int box_rows, bool f32) {
  return make_tmap(ctx, map, base, rows, cols, ld, box_cols, box_rows, f32);
}

int gemm_tc_init(DsContext* ctx) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DS_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  DS_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, DS_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  ctx->encode_tiled = fn;
  return DS_OK;
}

int gemm_tc_launch(DsContext* ctx, const GemmDesc& g, cudaStream_t s) {
  DS_CHECK(g.a_dtype == DT_BF16, DS_ERR_INVALID, "gemm_tc: A/W must be bf16");
  DS_CHECK(g.M > 0 && g.N > 0 && g.K > 0, DS_ERR_INVALID, "gemm_tc: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  DS_CHECK((g.lda % 8) == 0 && (g.ldw % 8) == 0, DS_ERR_INVALID, "gemm_tc: lda/ldw must be multiples of 8 (16 B rows)");
  DS_CHECK((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.W) & 15) == 0, DS_ERR_INVALID,
           "gemm_tc: A/W must be 16-byte aligned");
  switch (g.mode) {
    case GEMM_LNMOD:
      DS_CHECK(g.N == 64 && g.out_dtype == DT_BF16, DS_ERR_INVALID, "gemm_tc LNMOD: N must be 64, bf16 out");
      return launch_cfg<64, GEMM_LNMOD, true>(ctx, g, s);
    case GEMM_RESGATE:
      DS_CHECK((g.N == 64 || g.N == 256) && g.out_dtype == DT_F32, DS_ERR_INVALID, "gemm_tc RESGATE: N in {64,256}, fp32 out");
      DS_CHECK((g.ldo % 4) == 0 && (reinterpret_cast<uintptr_t>(g.out) & 15) == 0, DS_ERR_INVALID, "gemm_tc RESGATE: out alignment");
      if (g.N == 64) return launch_cfg<64, GEMM_RESGATE, true>(ctx, g, s);
      return launch_cfg<256, GEMM_RESGATE, true>(ctx, g, s);
    case GEMM_EHEAD:
      DS_CHECK(g.N == 64 && g.wc2 && g.wdir, DS_ERR_INVALID, "gemm_tc EHEAD: N must be 64");
      return launch_cfg<64, GEMM_EHEAD, false>(ctx, g, s);
    case GEMM_COORD:
      DS_CHECK(g.N == 256, DS_ERR_INVALID, "gemm_tc COORD: N must be 256");
      return launch_cfg<256, GEMM_COORD, false>(ctx, g, s);
    default:
      break;
  }
  // TMA stores clip the inner dimension at 16-byte granularity (measured): N must be a multiple of 8 bf16
  const bool tma_out = g.out_dtype == DT_BF16 && g.N >= 64 && (g.N % 8) == 0 && (g.ldo % 8) == 0 &&
                       (reinterpret_cast<uintptr_t>(g.out) & 15) == 0;
  DS_CHECK(g.split_n == 0 || (tma_out && g.N > 128), DS_ERR_INVALID, "gemm_tc: split output is implemented for the bf16 TMA-store path with 256-column tiles");
  if (g.N <= 16) return launch_cfg<16, GEMM_STORE, false>(ctx, g, s);
  if (g.N <= 32) return launch_cfg<32, GEMM_STORE, false>(ctx, g, s);
  if (g.N <= 64) return tma_out ? launch_cfg<64, GEMM_STORE, true>(ctx, g, s) : launch_cfg<64, GEMM_STORE, false>(ctx, g, s);
  if (g.N <= 128) return tma_out ? launch_cfg<128, GEMM_STORE, true>(ctx, g, s) : launch_cfg<128, GEMM_STORE, false>(ctx, g, s);
  return tma_out ? launch_cfg<256, GEMM_STORE, true>(ctx, g, s) : launch_cfg<256, GEMM_STORE, false>(ctx, g, s);
}
