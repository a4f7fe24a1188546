// Fused edge-stream update of one EquivariantMixBlock (models/dmt.py:156-157,165-169), one kernel per block:
//
//   e1 = modulate(LayerNorm(e_in + eg1 * (P[i] + P[j] + b_n2e)), esh2, esc2)       P = node2edge_lin applied per atom
//   e  = e1 + eg2 * (ff_linear4(SiLU(ff_linear3(e1))))                              64 -> 128 -> 64
//
// replacing k_edge_update1 + the ff3 GEMM + the ff4 RESGATE GEMM (three launches, 313 MB of HBM traffic per block)
// by ONE pass over the pair rows (104 MB: e in, e out, bf16 operand copy out).  Both weight matrices (16 KB each) stay
// resident in shared memory; per 128-row tile the operand of the first MMA is BUILT in shared memory by the producer
// warps (fp32 LayerNorm statistics by half-warp shuffles), the SiLU'd hidden tile goes TMEM -> registers -> swizzled
// shared memory -> second MMA without touching HBM, and the residual base e1 waits in shared memory (fp32).
//
// Persistent CTA (one per SM), 18 warps, everything double-buffered on tile parity s = it & 1:
//   warps 0-3    epilogue 2: e = e1 + gate * (acc2 + b4) -> global (fp32 stream + bf16 copy into the [dist | e] operand)
//   warps 4-7    epilogue 1: SiLU(acc1 + b3) -> bf16 -> A2[s] (SWIZZLE_128B K-major, two k-blocks)
//   warp  8      tcgen05.mma issuer + TMEM owner  (MMA1(it+1) is issued before MMA2(it): the tensor pipe never waits
//                for the SiLU epilogue of the same tile)
//   warp  9      TMA: ff_linear3 / ff_linear4 weights, once
//   warps 10-17  operand builders: 16 rows per warp and tile, half a warp per row, the loads of 8 rows in flight
#include "kernels.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int TM = 128;                       // pair rows per tile = UMMA M
constexpr int kBuilders = 8;
constexpr int kThreads = (4 + 4 + 2 + kBuilders) * 32;
constexpr int kA1 = TM * 64 * 2;              // 16 KB  bf16 [128 x 64], one k-block
constexpr int kA2 = TM * 128 * 2;             // 32 KB  bf16 [128 x 128], two k-blocks
constexpr int kE1Row = 272;                   // fp32 residual rows padded 256 -> 272 B: row-per-thread float4 reads are conflict-free
constexpr int kE1 = TM * kE1Row;              // 34 KB
constexpr int kW3 = 128 * 64 * 2;             // 16 KB  ff_linear3 [128, 64]
constexpr int kW4 = 64 * 128 * 2;             // 16 KB  ff_linear4 [64, 128] as two k-blocks of [64 x 64]
constexpr int kSmem = 2 * kA1 + 2 * kA2 + kW3 + kW4 + 2 * kE1 + 1024 /*biases*/ + 256 /*barriers*/;
constexpr uint32_t kAcc1Cols = 128, kAcc2Cols = 64, kStageCols = 192;

struct EdgeFfnArgs {
  float* e;                  // [Mp,64] fp32 stream, updated in place
  bf16* xe;                  // bf16 copy, row stride ldx (the e half of the [dist | e] operand)
  int ldx;
  const float* pn;           // [Mn,64] hoisted node2edge_lin (no bias)
  const float* n2e_b;        // [64]
  const float* ada;          // adaLN table pre-offset to block + ADA_EDGE; row stride ADA_LD
  const uint32_t* pair_info; // mol << 12 | i << 6 | j
  const int2* pair_rows;     // atom rows of (i, j)
  const float* b3;           // [128] (halved together with W3: SiLU(x) = h + h tanh(h))
  const float* b4;           // [64]
  int Mp;
};

__device__ __forceinline__ float half_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __launch_bounds__(kThreads, 1)
edge_ffn_kernel(const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmW4, EdgeFfnArgs a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smA1 = smem;                                  // [2][16 KB]
  uint8_t* smA2 = smA1 + 2 * kA1;                        // [2][32 KB]
  uint8_t* smW3 = smA2 + 2 * kA2;                        // 16 KB
  uint8_t* smW4 = smW3 + kW3;                            // 16 KB
  uint8_t* smE1 = smW4 + kW4;                            // [2][34 KB]
  float* sb3 = reinterpret_cast<float*>(smE1 + 2 * kE1); // [128]
  float* sb4 = sb3 + 128;                                // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb3 + 256);
  uint64_t* a1_full = bars;        // [2] builders -> MMA (and epilogue 2: E1 tile written)
  uint64_t* a1_empty = bars + 2;   // [2] MMA1 done reading A1
  uint64_t* t1_full = bars + 4;    // [2] MMA1 -> epilogue 1
  uint64_t* t1_empty = bars + 6;   // [2] epilogue 1 drained acc1
  uint64_t* a2_full = bars + 8;    // [2] epilogue 1 -> MMA2
  uint64_t* a2_empty = bars + 10;  // [2] MMA2 done reading A2
  uint64_t* t2_full = bars + 12;   // [2] MMA2 -> epilogue 2
  uint64_t* t2_empty = bars + 14;  // [2] epilogue 2 drained acc2
  uint64_t* e1_empty = bars + 16;  // [2] epilogue 2 done with the fp32 residual tile
  uint64_t* w_full = bars + 18;    // weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.Mp + TM - 1) / TM;
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 9 && lane == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();
    ptx::prefetch_tmap(&tmW3);
    ptx::prefetch_tmap(&tmW4);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a1_full[i], kBuilders);
      ptx::mbar_init(&a1_empty[i], 1);
      ptx::mbar_init(&t1_full[i], 1);
      ptx::mbar_init(&t1_empty[i], 128);
      ptx::mbar_init(&a2_full[i], 4);
      ptx::mbar_init(&a2_empty[i], 1);
      ptx::mbar_init(&t2_full[i], 1);
      ptx::mbar_init(&t2_empty[i], 128);
      ptx::mbar_init(&e1_empty[i], 128);
    }
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 8) ptx::tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < 128) sb3[threadIdx.x] = a.b3[threadIdx.x];
  else if (threadIdx.x < 192) sb4[threadIdx.x - 128] = a.b4[threadIdx.x - 128];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 9) {
    // ===================== weights: L2 -> shared, once =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(w_full, kW3 + kW4);
      ptx::tma_load_2d(smW3, &tmW3, w_full, 0, 0);               // [128 rows x 64 K]
      ptx::tma_load_2d(smW4, &tmW4, w_full, 0, 0);               // k-block 0: [64 rows x K 0..63]
      ptx::tma_load_2d(smW4 + kW4 / 2, &tmW4, w_full, 64, 0);    // k-block 1
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(TM, 128), idesc2 = ptx::umma_idesc_bf16(TM, 64);
      ptx::mbar_wait(w_full, 0);
      ptx::tc_fence_after();
      auto mma1 = [&](int it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&t1_empty[s], ph ^ 1);
        ptx::mbar_wait(&a1_full[s], ph);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + s * kStageCols;
        const uint32_t a_addr = ptx::smem_u32(smA1 + s * kA1), w_addr = ptx::smem_u32(smW3);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_bf16(d, ptx::umma_smem_desc_sw128(a_addr + k * 32), ptx::umma_smem_desc_sw128(w_addr + k * 32), idesc1, k ? 1u : 0u);
        ptx::umma_commit(&a1_empty[s]);
        ptx::umma_commit(&t1_full[s]);
      };
      auto mma2 = [&](int it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&t2_empty[s], ph ^ 1);
        ptx::mbar_wait(&a2_full[s], ph);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + s * kStageCols + kAcc1Cols;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t a_addr = ptx::smem_u32(smA2 + s * kA2 + kb * (kA2 / 2)), w_addr = ptx::smem_u32(smW4 + kb * (kW4 / 2));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(d, ptx::umma_smem_desc_sw128(a_addr + k * 32), ptx::umma_smem_desc_sw128(w_addr + k * 32), idesc2,
                           (kb | k) ? 1u : 0u);
        }
        ptx::umma_commit(&a2_empty[s]);
        ptx::umma_commit(&t2_full[s]);
      };
      if (my_tiles > 0) mma1(0);
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) mma1(it + 1);
        mma2(it);
      }
    }
  } else if (warp >= 10) {
    // ===================== builders: e1 = mod(LN(e + g1 (P_i + P_j + b))) -> A1 (bf16, swizzled) + E1 (fp32) =====================
    const int bw = warp - 10, hw = lane >> 4, hl = lane & 15, c0 = hl * 4;
    const float4 bb = *reinterpret_cast<const float4*>(a.n2e_b + c0);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      uint8_t* A1 = smA1 + s * kA1;
      uint8_t* E1 = smE1 + s * kE1;
      ptx::mbar_wait(&a1_empty[s], ph ^ 1);
      ptx::mbar_wait(&e1_empty[s], ph ^ 1);
#pragma unroll 1
      for (int batch = 0; batch < 2; ++batch) {
        // rows of this half-warp in the batch: lr = bw*16 + batch*8 + 2*q + hw, q = 0..3; all loads first
        float4 ev[4], pi[4], pj[4];
        uint32_t mol[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int lr = bw * 16 + batch * 8 + 2 * q + hw;
          const int p = min(tile * TM + lr, a.Mp - 1);
          const int2 rows = __ldg(a.pair_rows + p);
          mol[q] = __ldg(a.pair_info + p) >> 12;
          ev[q] = *reinterpret_cast<const float4*>(a.e + static_cast<size_t>(p) * 64 + c0);
          pi[q] = __ldg(reinterpret_cast<const float4*>(a.pn + static_cast<size_t>(rows.x) * 64 + c0));
          pj[q] = __ldg(reinterpret_cast<const float4*>(a.pn + static_cast<size_t>(rows.y) * 64 + c0));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int lr = bw * 16 + batch * 8 + 2 * q + hw;
          const float* ar = a.ada + static_cast<size_t>(mol[q]) * ADA_LD;
          const float4 g = __ldg(reinterpret_cast<const float4*>(ar + 128 + c0));
          float v0 = ev[q].x + g.x * ((pi[q].x + pj[q].x) + bb.x);
          float v1 = ev[q].y + g.y * ((pi[q].y + pj[q].y) + bb.y);
          float v2 = ev[q].z + g.z * ((pi[q].z + pj[q].z) + bb.z);
          float v3 = ev[q].w + g.w * ((pi[q].w + pj[q].w) + bb.w);
          const float mean = half_sum16((v0 + v1) + (v2 + v3)) * (1.0f / 64.0f);
          v0 -= mean; v1 -= mean; v2 -= mean; v3 -= mean;
          const float var = half_sum16((v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3)) * (1.0f / 64.0f);
          const float is = rsqrtf(var + 1e-6f);
          const float4 sh = __ldg(reinterpret_cast<const float4*>(ar + 192 + c0));
          const float4 sc = __ldg(reinterpret_cast<const float4*>(ar + 256 + c0));
          v0 = (v0 * is) * (1.0f + sc.x) + sh.x;
          v1 = (v1 * is) * (1.0f + sc.y) + sh.y;
          v2 = (v2 * is) * (1.0f + sc.z) + sh.z;
          v3 = (v3 * is) * (1.0f + sc.w) + sh.w;
          *reinterpret_cast<float4*>(E1 + lr * kE1Row + c0 * 4) = make_float4(v0, v1, v2, v3);
          *reinterpret_cast<uint2*>(A1 + lr * 128 + (((hl >> 1) ^ (lr & 7)) << 4) + (hl & 1) * 8) =
              make_uint2(pack2(v0, v1), pack2(v2, v3));
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&a1_full[s]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue 1: SiLU(acc1 + b3) -> A2 (bf16, swizzled, two k-blocks) =====================
    const int wq = warp & 3, r = wq * 32 + lane;
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      uint8_t* A2 = smA2 + s * kA2;
      ptx::mbar_wait(&a2_empty[s], ph ^ 1);
      ptx::mbar_wait(&t1_full[s], ph);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + s * kStageCols;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld32_sync(t_addr + c, v);
        uint8_t* kbp = A2 + (c >> 6) * (kA2 / 2) + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = act_silu_half<true>(__uint_as_float(v[q * 8 + i]) + sb3[c + q * 8 + i]);
          const int chunk = ((c & 63) >> 3) + q;
          *reinterpret_cast<uint4*>(kbp + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&t1_empty[s]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&a2_full[s]);
    }
  } else {
    // ===================== epilogue 2: e = e1 + gate * (acc2 + b4) -> global =====================
    const int wq = warp, r = wq * 32 + lane;
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int p = tile * TM + r;
      const bool ok = p < a.Mp;
      const uint32_t mol = ok ? (__ldg(a.pair_info + p) >> 12) : 0u;
      const float* gr = a.ada + static_cast<size_t>(mol) * ADA_LD + 320;
      const uint8_t* E1 = smE1 + s * kE1 + r * kE1Row;
      ptx::mbar_wait(&a1_full[s], ph);          // residual tile written (long done; orders the generic-proxy reads)
      ptx::mbar_wait(&t2_full[s], ph);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + s * kStageCols + kAcc1Cols;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld32_sync(t_addr + c, v);
        if (ok) {
          float* eo = a.e + static_cast<size_t>(p) * 64 + c;
          bf16* xo = a.xe + static_cast<size_t>(p) * a.ldx + c;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 r0 = *reinterpret_cast<const float4*>(E1 + (c + q * 8) * 4);
            const float4 r1 = *reinterpret_cast<const float4*>(E1 + (c + q * 8 + 4) * 4);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gr + c + q * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gr + c + q * 8 + 4));
            float o[8];
            o[0] = r0.x + g0.x * (__uint_as_float(v[q * 8 + 0]) + sb4[c + q * 8 + 0]);
            o[1] = r0.y + g0.y * (__uint_as_float(v[q * 8 + 1]) + sb4[c + q * 8 + 1]);
            o[2] = r0.z + g0.z * (__uint_as_float(v[q * 8 + 2]) + sb4[c + q * 8 + 2]);
            o[3] = r0.w + g0.w * (__uint_as_float(v[q * 8 + 3]) + sb4[c + q * 8 + 3]);
            o[4] = r1.x + g1.x * (__uint_as_float(v[q * 8 + 4]) + sb4[c + q * 8 + 4]);
            o[5] = r1.y + g1.y * (__uint_as_float(v[q * 8 + 5]) + sb4[c + q * 8 + 5]);
            o[6] = r1.z + g1.z * (__uint_as_float(v[q * 8 + 6]) + sb4[c + q * 8 + 6]);
            o[7] = r1.w + g1.w * (__uint_as_float(v[q * 8 + 7]) + sb4[c + q * 8 + 7]);
            *reinterpret_cast<float4*>(eo + q * 8) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(eo + q * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
            *reinterpret_cast<uint4*>(xo + q * 8) = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&t2_empty[s]);
      ptx::mbar_arrive(&e1_empty[s]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int edge_ffn_launch(DsContext* ctx, const Plan& plan, float* e, void* xe, int ldx, const float* pn, const float* n2e_b,
                    const float* ada_l, const void* w3, const float* b3, const void* w4, const float* b4, cudaStream_t s) {
  if (plan.Mp <= 0) return DS_OK;
  DS_CHECK(plan.pair_rows != nullptr, DS_ERR_INVALID, "edge_ffn: plan has no pair-row table");
  static bool attr_set = false;
  if (!attr_set) {
    DS_CUDA_CHECK(cudaFuncSetAttribute(edge_ffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  CUtensorMap tmW3, tmW4;
  DS_TRY(ds_make_tmap_2d(ctx, &tmW3, w3, 128, 64, 64, 64, 128, false));
  DS_TRY(ds_make_tmap_2d(ctx, &tmW4, w4, 64, 128, 128, 64, 64, false));
  EdgeFfnArgs a;
  a.e = e;
  a.xe = reinterpret_cast<bf16*>(xe);
  a.ldx = ldx;
  a.pn = pn;
  a.n2e_b = n2e_b;
  a.ada = ada_l + ADA_EDGE;
  a.pair_info = plan.pair_info;
  a.pair_rows = plan.pair_rows;
  a.b3 = b3;
  a.b4 = b4;
  a.Mp = plan.Mp;
  const int tiles = (plan.Mp + TM - 1) / TM;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  ds_launch(edge_ffn_kernel, dim3(grid), dim3(kThreads), kSmem, s, tmW3, tmW4, a);
  DS_CUDA_CHECK(cudaGetLastError());
  ctx->launch_count++;
  return DS_OK;
}
