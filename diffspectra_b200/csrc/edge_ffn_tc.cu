// Fused edge-stream update of one EquivariantMixBlock (models/dmt.py:156-157,165-169), one kernel per block:
//
//   e1 = modulate(LayerNorm(e_in + eg1 * (P[i] + P[j] + b_n2e)), esh2, esc2)       P = node2edge_lin applied per atom
//   e  = e1 + eg2 * (ff_linear4(SiLU(ff_linear3(e1))))                              64 -> 128 -> 64
//
// replacing k_edge_update1 + the ff3 GEMM + the ff4 RESGATE GEMM (three launches, 313 MB of HBM traffic per block)
// by ONE pass over the pair rows (104 MB: e in, e out, bf16 operand copy out).
//
// Persistent CTA (one per SM), 14 warps.  Three WARP-GROUPS each own one 128-row tile at a time and walk it through
// the whole chain, one THREAD per pair row:
//   stage the tile's e rows (32 KB, coalesced) -> each thread keeps its row in registers (64 fp32): residual sum,
//   LayerNorm and modulate are thread-local (no shuffles) -> bf16 row into the SWIZZLE_128B operand tile -> MMA1 ->
//   tcgen05.ld, SiLU, bf16 row into the second operand tile (same shared memory) -> MMA2 (accumulator re-uses the
//   TMEM columns of MMA1) -> tcgen05.ld, gated residual on the registers kept since step one -> staged, coalesced
//   stores of the fp32 stream and of its bf16 copy.
// The bf16 copy of the updated row, staged in SWIZZLE_128B K-major form for its coalesced store, doubles as the operand of a
// THIRD MMA: the block's skip projection into the edge heads, edge_l(e) [64 -> 16] (models/dmt.py:387-388), which used
// to be a separate GEMM launch over all pairs.
// Three tiles are in flight per SM and the only cross-warp hand-offs are the six mbarriers per group shared with the
// single MMA-issuing thread (warp 12), which polls the groups round-robin.  Warp 13 loads the two 16 KB weight
// matrices once.  32 KB of shared memory per group serve, in turn, as e staging, A1, A2 and output staging.
#include "kernels.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int TM = 128;                       // pair rows per tile = UMMA M = threads per warp-group
constexpr int kGroups = 3;
constexpr int kThreads = (kGroups * 4 + 2) * 32;
constexpr int kBuf = 32 * 1024;               // per group: e staging / A1 (first 16 KB) / A2 / output staging
constexpr int kW3 = 128 * 64 * 2;             // ff_linear3 [128, 64] bf16
constexpr int kW4 = 64 * 128 * 2;             // ff_linear4 [64, 128] bf16 as two k-blocks of [64 x 64]
constexpr int kWs = 16 * 64 * 2;              // edge_l [16, 64] bf16 (skip projection into the edge heads)
constexpr int kSmem = kGroups * kBuf + kW3 + kW4 + kWs + 1024 /*biases*/ + 256 /*barriers*/;
constexpr uint32_t kGroupCols = 128;          // TMEM columns per group: acc1 [0,128), acc2 re-uses [0,64), acc3 (skip) [64,80)

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
  const float* bs;           // [16] bias of the skip projection, or null: no skip output
  bf16* skip;                // skip output: 16 columns per pair row, row stride lds (elements)
  int lds;
  int Mp;
};

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }
// read-only 16-byte load the compiler may not sink next to its first use: a batch of these stays a batch in flight
__device__ __forceinline__ float4 ldg128f(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// kCoop: the pre-LayerNorm row x = e + eg1 * (P[i] + P[j] + b) is built while the tile is STAGED, by the threads that
// own consecutive 16-byte chunks of a row (16 lanes read one 256-byte P row), instead of by the row's own thread (32
// lanes reading 16 bytes each of 32 different P rows: 32 L1 wavefronts per load instruction, 32 such loads per thread).
template <bool kCoop>
__global__ void __launch_bounds__(kThreads, 1)
edge_ffn_kernel(const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmW4,
                const __grid_constant__ CUtensorMap tmWs, EdgeFfnArgs a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smBuf = smem;                                   // [kGroups][32 KB]
  uint8_t* smW3 = smBuf + kGroups * kBuf;                  // 16 KB
  uint8_t* smW4 = smW3 + kW3;                              // 16 KB
  uint8_t* smWs = smW4 + kW4;                              // 2 KB
  float* sb3 = reinterpret_cast<float*>(smWs + kWs);       // [128]
  float* sb4 = sb3 + 128;                                  // [64]
  float* sbn = sb4 + 64;                                   // [64] node2edge_lin bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb3 + 256);
  uint64_t* a1_full = bars;                  // [3] group -> MMA: A1 written (128 arrivals)
  uint64_t* t1_full = bars + kGroups;        // [3] MMA1 done
  uint64_t* a2_full = bars + 2 * kGroups;    // [3] group -> MMA: A2 written, acc1 drained
  uint64_t* t2_full = bars + 3 * kGroups;    // [3] MMA2 done
  uint64_t* a3_full = bars + 4 * kGroups;    // [3] group -> MMA: bf16 row tile staged (operand of the skip projection)
  uint64_t* t3_full = bars + 5 * kGroups;    // [3] MMA3 done
  uint64_t* w_full = bars + 6 * kGroups;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 * kGroups + 1);
  const bool has_skip = a.bs != nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.Mp + TM - 1) / TM;
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 13 && lane == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();
    ptx::prefetch_tmap(&tmW3);
    ptx::prefetch_tmap(&tmW4);
    ptx::prefetch_tmap(&tmWs);
    for (int i = 0; i < kGroups; ++i) {
      ptx::mbar_init(&a1_full[i], 128);
      ptx::mbar_init(&t1_full[i], 1);
      ptx::mbar_init(&a2_full[i], 128);
      ptx::mbar_init(&t2_full[i], 1);
      ptx::mbar_init(&a3_full[i], 128);
      ptx::mbar_init(&t3_full[i], 1);
    }
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 12) ptx::tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < 128) sb3[threadIdx.x] = a.b3[threadIdx.x];
  else if (threadIdx.x < 192) sb4[threadIdx.x - 128] = a.b4[threadIdx.x - 128];
  else if (threadIdx.x < 256) sbn[threadIdx.x - 192] = a.n2e_b[threadIdx.x - 192];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 13) {
    // ===================== weights: L2 -> shared, once =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(w_full, kW3 + kW4 + (has_skip ? kWs : 0));
      if (has_skip) ptx::tma_load_2d(smWs, &tmWs, w_full, 0, 0);      // [16 rows x 64 K]
      ptx::tma_load_2d(smW3, &tmW3, w_full, 0, 0);               // [128 rows x 64 K]
      ptx::tma_load_2d(smW4, &tmW4, w_full, 0, 0);               // k-block 0: [64 rows x K 0..63]
      ptx::tma_load_2d(smW4 + kW4 / 2, &tmW4, w_full, 64, 0);    // k-block 1
    }
  } else if (warp == 12) {
    // ===================== MMA issuer: serves the groups in whatever order their operands become ready =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(TM, 128), idesc2 = ptx::umma_idesc_bf16(TM, 64), idesc3 = ptx::umma_idesc_bf16(TM, 16);
      ptx::mbar_wait(w_full, 0);
      ptx::tc_fence_after();
      int left[kGroups], stage[kGroups];
      uint32_t ph[kGroups];
      int open = 0;
      for (int g = 0; g < kGroups; ++g) {
        left[g] = (my_tiles - g + kGroups - 1) / kGroups;
        stage[g] = 0;
        ph[g] = 0;
        if (left[g] > 0) ++open;
      }
      while (open > 0) {
        for (int g = 0; g < kGroups; ++g) {
          if (left[g] <= 0) continue;
          const uint32_t d = tmem_base + g * kGroupCols;
          const uint32_t buf = ptx::smem_u32(smBuf + g * kBuf);
          if (stage[g] == 0) {
            if (!ptx::mbar_try_wait(&a1_full[g], ph[g])) continue;
            ptx::tc_fence_after();
            const uint32_t w_addr = ptx::smem_u32(smW3);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(d, ptx::umma_smem_desc_sw128(buf + k * 32), ptx::umma_smem_desc_sw128(w_addr + k * 32), idesc1, k ? 1u : 0u);
            ptx::umma_commit(&t1_full[g]);
            stage[g] = 1;
          } else if (stage[g] == 2) {
            // skip projection: the staged bf16 rows [128 x 64] x edge_l^T [16 x 64] -> 16 accumulator columns
            if (!ptx::mbar_try_wait(&a3_full[g], ph[g])) continue;
            ptx::tc_fence_after();
            const uint32_t w_addr = ptx::smem_u32(smWs);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16(d + 64, ptx::umma_smem_desc_sw128(buf + k * 32), ptx::umma_smem_desc_sw128(w_addr + k * 32), idesc3, k ? 1u : 0u);
            ptx::umma_commit(&t3_full[g]);
            stage[g] = 0;
            ph[g] ^= 1;
            if (--left[g] == 0) --open;
          } else {
            if (!ptx::mbar_try_wait(&a2_full[g], ph[g])) continue;
            ptx::tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint32_t a_addr = buf + kb * (kBuf / 2), w_addr = ptx::smem_u32(smW4 + kb * (kW4 / 2));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_bf16(d, ptx::umma_smem_desc_sw128(a_addr + k * 32), ptx::umma_smem_desc_sw128(w_addr + k * 32), idesc2,
                               (kb | k) ? 1u : 0u);
            }
            ptx::umma_commit(&t2_full[g]);
            if (has_skip) {
              stage[g] = 2;
            } else {
              stage[g] = 0;
              ph[g] ^= 1;
              if (--left[g] == 0) --open;
            }
          }
        }
      }
    }
  } else {
    // ===================== warp-groups: one thread per pair row, the whole chain of a tile =====================
    const int g = warp >> 2, wq = warp & 3, r = wq * 32 + lane;     // r = row inside the tile = TMEM lane
    uint8_t* buf = smBuf + g * kBuf;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + g * kGroupCols;
    uint32_t ph = 0;
    for (int k = g; k < my_tiles; k += kGroups, ph ^= 1) {
      const int tile = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
      const int p0 = tile * TM;
      const int p = p0 + r;
      const bool ok = p < a.Mp;
      const int pc = ok ? p : a.Mp - 1;
      const int2 rows = __ldg(a.pair_rows + pc);
      const float* ar = a.ada + static_cast<size_t>(__ldg(a.pair_info + pc) >> 12) * ADA_LD;
      if constexpr (kCoop) {     // shift | scale | gate2 of the row's molecule (768 bytes) are needed two and four steps from now
#pragma unroll
        for (int i = 0; i < 6; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(ar + 192 + i * 32));
      }
      // ---- 1. stage the tile's e rows: 2048 float4, coalesced, 16-byte chunks XOR-swizzled by the row
      if constexpr (kCoop) {
        const float4* src = reinterpret_cast<const float4*>(a.e + static_cast<size_t>(p0) * 64);
        const int nrow = min(a.Mp - p0, TM);
        const int ch = r & 15;                       // this thread's chunk of every row it touches
        const float4 bn = *reinterpret_cast<const float4*>(sbn + ch * 4);
#pragma unroll
        for (int qb = 0; qb < 16; qb += 4) {         // four rows per batch: 16 independent 16-byte loads in flight
          float4 te[4], tpi[4], tpj[4], tg[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int row = (qb + u) * 8 + (r >> 4);
            const int prow = p0 + min(row, nrow - 1);
            const int2 rr = __ldg(a.pair_rows + prow);
            const float* arr = a.ada + static_cast<size_t>(__ldg(a.pair_info + prow) >> 12) * ADA_LD;
            te[u] = row < nrow ? src[row * 16 + ch] : make_float4(0.f, 0.f, 0.f, 0.f);
            tpi[u] = ldg128f(a.pn + static_cast<size_t>(rr.x) * 64 + ch * 4);
            tpj[u] = ldg128f(a.pn + static_cast<size_t>(rr.y) * 64 + ch * 4);
            tg[u] = ldg128f(arr + 128 + ch * 4);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int row = (qb + u) * 8 + (r >> 4);
            float4 x;
            x.x = te[u].x + tg[u].x * ((tpi[u].x + tpj[u].x) + bn.x);
            x.y = te[u].y + tg[u].y * ((tpi[u].y + tpj[u].y) + bn.y);
            x.z = te[u].z + tg[u].z * ((tpi[u].z + tpj[u].z) + bn.z);
            x.w = te[u].w + tg[u].w * ((tpi[u].w + tpj[u].w) + bn.w);
            *reinterpret_cast<float4*>(buf + row * 256 + ((ch ^ (row & 15)) << 4)) = x;
          }
        }
      } else {
        const float4* src = reinterpret_cast<const float4*>(a.e + static_cast<size_t>(p0) * 64);
        const int lim = (min(a.Mp - p0, TM)) * 16;
        float4 t[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int idx = q * 128 + r;
          t[q] = idx < lim ? src[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int idx = q * 128 + r, row = idx >> 4, ch = idx & 15;
          *reinterpret_cast<float4*>(buf + row * 256 + ((ch ^ (row & 15)) << 4)) = t[q];
        }
      }
      group_sync(g);
      // ---- 2. own row -> registers; residual sum, LayerNorm, modulate: all thread-local
      float v[64];
      float sum = 0.f;
      if constexpr (kCoop) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 ev = *reinterpret_cast<const float4*>(buf + r * 256 + ((c ^ (r & 15)) << 4));
          v[c * 4 + 0] = ev.x; v[c * 4 + 1] = ev.y; v[c * 4 + 2] = ev.z; v[c * 4 + 3] = ev.w;
          sum += (ev.x + ev.y) + (ev.z + ev.w);
        }
      } else
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {             // two halves of 32 channels: 24 independent loads in flight each
        float4 pi[8], pj[8], g1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          pi[c] = ldg128f(a.pn + static_cast<size_t>(rows.x) * 64 + (hf * 8 + c) * 4);
          pj[c] = ldg128f(a.pn + static_cast<size_t>(rows.y) * 64 + (hf * 8 + c) * 4);
          g1[c] = ldg128f(ar + 128 + (hf * 8 + c) * 4);
        }
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const int c = hf * 8 + c8;
          const float4 ev = *reinterpret_cast<const float4*>(buf + r * 256 + ((c ^ (r & 15)) << 4));
          v[c * 4 + 0] = ev.x + g1[c8].x * ((pi[c8].x + pj[c8].x) + sbn[c * 4 + 0]);
          v[c * 4 + 1] = ev.y + g1[c8].y * ((pi[c8].y + pj[c8].y) + sbn[c * 4 + 1]);
          v[c * 4 + 2] = ev.z + g1[c8].z * ((pi[c8].z + pj[c8].z) + sbn[c * 4 + 2]);
          v[c * 4 + 3] = ev.w + g1[c8].w * ((pi[c8].w + pj[c8].w) + sbn[c * 4 + 3]);
          sum += (v[c * 4 + 0] + v[c * 4 + 1]) + (v[c * 4 + 2] + v[c * 4 + 3]);
        }
      }
      const float mean = sum * (1.0f / 64.0f);
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        v[c] -= mean;
        sq = fmaf(v[c], v[c], sq);
      }
      const float is = rsqrtf(sq * (1.0f / 64.0f) + 1e-6f);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float4 sh[8], sc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          sh[c] = ldg128f(ar + 192 + (hf * 8 + c) * 4);
          sc[c] = ldg128f(ar + 256 + (hf * 8 + c) * 4);
        }
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const int c = hf * 8 + c8;
          v[c * 4 + 0] = (v[c * 4 + 0] * is) * (1.0f + sc[c8].x) + sh[c8].x;
          v[c * 4 + 1] = (v[c * 4 + 1] * is) * (1.0f + sc[c8].y) + sh[c8].y;
          v[c * 4 + 2] = (v[c * 4 + 2] * is) * (1.0f + sc[c8].z) + sh[c8].z;
          v[c * 4 + 3] = (v[c * 4 + 3] * is) * (1.0f + sc[c8].w) + sh[c8].w;
        }
      }
      group_sync(g);                               // every thread has read its staged row: the buffer becomes A1
      // ---- 3. A1 row (bf16, SWIZZLE_128B K-major: row r, 16-byte chunk c at r*128 + ((c ^ (r & 7)) << 4))
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(buf + r * 128 + ((c ^ (r & 7)) << 4)) =
            make_uint4(pack2(v[c * 8 + 0], v[c * 8 + 1]), pack2(v[c * 8 + 2], v[c * 8 + 3]), pack2(v[c * 8 + 4], v[c * 8 + 5]),
                       pack2(v[c * 8 + 6], v[c * 8 + 7]));
      ptx::tc_fence_before();                      // orders the previous tile's tcgen05.ld of acc2 before the next MMA1
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&a1_full[g]);
      // ---- 4. SiLU(acc1 + b3) -> A2 (two k-blocks of 16 KB; MMA1 has finished reading A1 once t1_full fires)
      ptx::mbar_wait(&t1_full[g], ph);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t acc[32];
        ptx::tmem_ld32_sync(t_addr + c, acc);
        uint8_t* kbp = buf + (c >> 6) * (kBuf / 2) + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = act_silu_half<true>(__uint_as_float(acc[q * 8 + i]) + sb3[c + q * 8 + i]);
          const int chunk = ((c & 63) >> 3) + q;
          *reinterpret_cast<uint4*>(kbp + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
        }
      }
      ptx::tc_fence_before();                      // acc1 drained: MMA2 may overwrite its first 64 columns
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&a2_full[g]);
      // ---- 5. e = e1 + eg2 * (acc2 + b4); staged, coalesced stores
      ptx::mbar_wait(&t2_full[g], ph);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 64; c += 32) {            // fully unrolled: v[] must keep static indices (registers)
        float4 g2[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) g2[q] = ldg128f(ar + 320 + c + q * 4);
        uint32_t acc[32];
        ptx::tmem_ld32_sync(t_addr + c, acc);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[c + q * 4 + 0] += g2[q].x * (__uint_as_float(acc[q * 4 + 0]) + sb4[c + q * 4 + 0]);
          v[c + q * 4 + 1] += g2[q].y * (__uint_as_float(acc[q * 4 + 1]) + sb4[c + q * 4 + 1]);
          v[c + q * 4 + 2] += g2[q].z * (__uint_as_float(acc[q * 4 + 2]) + sb4[c + q * 4 + 2]);
          v[c + q * 4 + 3] += g2[q].w * (__uint_as_float(acc[q * 4 + 3]) + sb4[c + q * 4 + 3]);
        }
      }
      // MMA2 has finished reading A2 (t2_full): the buffer becomes the fp32 output staging tile
#pragma unroll
      for (int c = 0; c < 16; ++c)
        *reinterpret_cast<float4*>(buf + r * 256 + ((c ^ (r & 15)) << 4)) = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
      group_sync(g);
      {
        float4* dst = reinterpret_cast<float4*>(a.e + static_cast<size_t>(p0) * 64);
        const int lim = (min(a.Mp - p0, TM)) * 16;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int idx = q * 128 + r, row = idx >> 4, ch = idx & 15;
          if (idx < lim) dst[idx] = *reinterpret_cast<const float4*>(buf + row * 256 + ((ch ^ (row & 15)) << 4));
        }
      }
      group_sync(g);
      // bf16 copy: 128-byte rows, staged the same way
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(buf + r * 128 + ((c ^ (r & 7)) << 4)) =
            make_uint4(pack2(v[c * 8 + 0], v[c * 8 + 1]), pack2(v[c * 8 + 2], v[c * 8 + 3]), pack2(v[c * 8 + 4], v[c * 8 + 5]),
                       pack2(v[c * 8 + 6], v[c * 8 + 7]));
      if (has_skip) {                              // the staged rows are also the operand of the skip projection
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&a3_full[g]);
      }
      group_sync(g);
      {
        const int nrow = min(a.Mp - p0, TM);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int idx = q * 128 + r, row = idx >> 3, ch = idx & 7;
          if (row < nrow)
            *reinterpret_cast<uint4*>(a.xe + static_cast<size_t>(p0 + row) * a.ldx + ch * 8) =
                *reinterpret_cast<const uint4*>(buf + row * 128 + ((ch ^ (row & 7)) << 4));
        }
      }
      if (has_skip) {
        ptx::mbar_wait(&t3_full[g], ph);           // MMA3 has also finished reading the staged rows
        ptx::tc_fence_after();
        uint32_t acc[16];
        ptx::tmem_ld16_sync(t_addr + 64, acc);
        if (ok) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = pack2(__uint_as_float(acc[2 * i]) + __ldg(a.bs + 2 * i), __uint_as_float(acc[2 * i + 1]) + __ldg(a.bs + 2 * i + 1));
          uint4* dst = reinterpret_cast<uint4*>(a.skip + static_cast<size_t>(p) * a.lds);
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      group_sync(g);                               // staging reads done before the next tile overwrites the buffer
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}


}  // namespace

int edge_ffn_launch(DsContext* ctx, const Plan& plan, float* e, void* xe, int ldx, const float* pn, const float* n2e_b,
                    const float* ada_l, const void* w3, const float* b3, const void* w4, const float* b4, const void* ws,
                    const float* bs, void* skip, int lds, cudaStream_t s) {
  if (plan.Mp <= 0) return DS_OK;
  DS_CHECK(plan.pair_rows != nullptr, DS_ERR_INVALID, "edge_ffn: plan has no pair-row table");
  static bool attr_set[64] = {};            // the attribute is per device: one flag per device ordinal
  if (!attr_set[ctx->device & 63]) {
    DS_CUDA_CHECK(cudaFuncSetAttribute(edge_ffn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    DS_CUDA_CHECK(cudaFuncSetAttribute(edge_ffn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set[ctx->device & 63] = true;
  }
  CUtensorMap tmW3, tmW4, tmWs;
  DS_TRY(ds_make_tmap_2d(ctx, &tmW3, w3, 128, 64, 64, 64, 128, false));
  DS_TRY(ds_make_tmap_2d(ctx, &tmW4, w4, 64, 128, 128, 64, 64, false));
  tmWs = tmW3;
  if (ws != nullptr) {
    DS_CHECK(bs != nullptr && skip != nullptr && (lds % 8) == 0 && (reinterpret_cast<uintptr_t>(skip) & 15) == 0, DS_ERR_INVALID,
             "edge_ffn: skip output needs a bias, a 16-byte aligned destination and a row stride that is a multiple of 8");
    DS_TRY(ds_make_tmap_2d(ctx, &tmWs, ws, 16, 64, 64, 64, 16, false));
  }
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
  a.bs = ws != nullptr ? bs : nullptr;
  a.skip = reinterpret_cast<bf16*>(skip);
  a.lds = lds;
  a.Mp = plan.Mp;
  const int tiles = (plan.Mp + TM - 1) / TM;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  if (ctx->ffn_variant == 1) {
    ds_launch(edge_ffn_kernel<true>, dim3(grid), dim3(kThreads), kSmem, s, tmW3, tmW4, tmWs, a);
  } else {
    ds_launch(edge_ffn_kernel<false>, dim3(grid), dim3(kThreads), kSmem, s, tmW3, tmW4, tmWs, a);
  }
  DS_CUDA_CHECK(cudaGetLastError());
  ctx->launch_count++;
  return DS_OK;
}
