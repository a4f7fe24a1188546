// Fused SE(3)-equivariant coordinate head of one EquivariantMixBlock (models/dmt.py:37-60, MultiCondEquiUpdate):
//
//   per directed edge d = (r -> c):   y = input_lin([h_r | h_c | e | dist])            = A[r] + B[c] + G[pair]
//                                     z = modulate(LayerNorm(y), csh[mol], csc[mol])                  dmt.py:42-44
//                                     u = SiLU(coord_mlp.0 z + b)          256 x 256 tcgen05 MMA      dmt.py:45
//                                     w = mean(tanh(coord_mlp.2 u) * [1, adj2d, adjsp])              dmt.py:46-51
//
// A, B (per atom) and G (per unordered pair) come from the hoisted GEMMs; nothing per directed edge ever touches
// HBM except the 4-byte result: the LayerNorm'd operand tile is BUILT in shared memory, in the SWIZZLE_128B K-major
// layout the UMMA descriptor expects, by 16 producer warps (one warp per row, fp32 statistics by warp shuffles),
// consumed by tcgen05.mma into a double-buffered TMEM accumulator, and reduced to w by two epilogue warp-groups.
//
// Persistent CTA (one per SM), 26 warps:
//   warps 0-3 / 4-7   epilogue groups (TMEM accumulator stage 0 / 1, even / odd tiles)
//   warp 8            TMA producer for coord_mlp.0 (two 32 KB k-block slots, re-read from L2 for every tile)
//   warp 9            tcgen05.mma issuer + TMEM owner
//   warps 10-25       operand builders (8 rows of the 128-row tile each, the loads of 4 rows in flight at a time)
#include "kernels.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int TM = 128;                 // rows (directed edges) per tile = UMMA M
constexpr int TK = 256;                 // hidden dim = K = N
constexpr int KB = 64;                  // one 128-byte swizzle atom of bf16
constexpr int NKB = TK / KB;            // 4 k-blocks
constexpr int kBuilders = 16;
constexpr int kRowsPerBuilder = TM / kBuilders;   // 8
constexpr int kBatch = 4;                         // rows whose loads are in flight together per builder warp
constexpr int kThreads = (8 + 2 + kBuilders) * 32;
constexpr int kZBytes = TM * TK * 2;              // 64 KB operand tile (4 k-blocks of 16 KB)
constexpr int kZkb = TM * KB * 2;                 // 16 KB
constexpr int kWSlot = TK * KB * 2;               // 32 KB: coord_mlp.0[:, kb*64:(kb+1)*64]
constexpr int kSmem = 2 * kZBytes + 2 * kWSlot + 256 * 16 + 256;

struct CoordArgs {
  const bf16* ab;          // [Mn,512]  A = cols 0..255 (with input_lin.bias), B = cols 256..511
  const bf16* gp;          // [Mp,256]
  const float* ada;        // adaLN table pre-offset to the block; row stride ADA_LD
  const uint8_t* pflags;   // [Mp] adjacency bits
  const int4* dir_info;    // [Md] pair row, node row of r, node row of c, molecule
  const float* bias;       // coord_mlp.0 bias [256]
  const float* wc2;        // coord_mlp.2 [3,256]
  float* wdir;             // [Md]
  int Md;
};

__device__ __forceinline__ void load8_f32(const float* row, int lane, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(row + 4 * lane);
  const float4 b = *reinterpret_cast<const float4*>(row + 128 + 4 * lane);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void unpack8(const uint2 (&r)[2], float (&v)[8]) {
  v[0] = __uint_as_float(r[0].x << 16); v[1] = __uint_as_float(r[0].x & 0xffff0000u);
  v[2] = __uint_as_float(r[0].y << 16); v[3] = __uint_as_float(r[0].y & 0xffff0000u);
  v[4] = __uint_as_float(r[1].x << 16); v[5] = __uint_as_float(r[1].x & 0xffff0000u);
  v[6] = __uint_as_float(r[1].y << 16); v[7] = __uint_as_float(r[1].y & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// lane's 4+4 channels of tile row `lr` -> SWIZZLE_128B K-major operand tile (k-block kb = channel / 64)
__device__ __forceinline__ void store_z_row(uint8_t* ztile, int lr, int lane, const float (&v)[8]) {
  const int kb = lane >> 4;
  const int chunk = (lane & 15) >> 1;
  uint8_t* p = ztile + kb * kZkb + lr * 128 + ((chunk ^ (lr & 7)) << 4) + (lane & 1) * 8;
  *reinterpret_cast<uint2*>(p) = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
  *reinterpret_cast<uint2*>(p + 2 * kZkb) = make_uint2(pack2(v[4], v[5]), pack2(v[6], v[7]));
}

__global__ void __launch_bounds__(kThreads, 1)
coord_fused_kernel(const __grid_constant__ CUtensorMap tmW, CoordArgs a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smZ = smem;                              // [2][64 KB]
  uint8_t* smW = smem + 2 * kZBytes;                // [2][32 KB]
  float4* swc2 = reinterpret_cast<float4*>(smW + 2 * kWSlot);   // [256] (w0, w1, w2, bias)
  uint64_t* bars = reinterpret_cast<uint64_t*>(swc2 + 256);
  uint64_t* zfull = bars;          // [2] builders -> MMA
  uint64_t* zempty = bars + 2;     // [2] MMA -> builders
  uint64_t* wfull = bars + 4;      // [2] TMA -> MMA
  uint64_t* wempty = bars + 6;     // [2] MMA -> TMA
  uint64_t* tfull = bars + 8;      // [2] MMA -> epilogue
  uint64_t* tempty = bars + 10;    // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.Md + TM - 1) / TM;
  const int my_tiles = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 8 && lane == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();
    ptx::prefetch_tmap(&tmW);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&zfull[i], kBuilders);
      ptx::mbar_init(&zempty[i], 1);
      ptx::mbar_init(&wfull[i], 1);
      ptx::mbar_init(&wempty[i], 1);
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < 256)
    swc2[threadIdx.x] = make_float4(a.wc2[threadIdx.x], a.wc2[256 + threadIdx.x], a.wc2[512 + threadIdx.x], a.bias[threadIdx.x]);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 8) {
    // ===================== coord_mlp.0 k-blocks: L2 -> shared =====================
    if (lane == 0) {
      for (int it = 0; it < my_tiles; ++it)
        for (int kb = 0; kb < NKB; ++kb) {
          const int idx = it * NKB + kb, slot = idx & 1;
          ptx::mbar_wait(&wempty[slot], ((idx >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&wfull[slot], kWSlot);
          ptx::tma_load_2d(smW + slot * kWSlot, &tmW, &wfull[slot], kb * KB, 0);
        }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(TM, TK);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        ptx::mbar_wait(&tempty[s], ph ^ 1);
        ptx::mbar_wait(&zfull[s], ph);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(s * TK);
        for (int kb = 0; kb < NKB; ++kb) {
          const int idx = it * NKB + kb, slot = idx & 1;
          ptx::mbar_wait(&wfull[slot], (idx >> 1) & 1);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smZ + s * kZBytes + kb * kZkb);
          const uint32_t w_addr = ptx::smem_u32(smW + slot * kWSlot);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k)
            ptx::umma_bf16(d_tmem, ptx::umma_smem_desc_sw128(a_addr + k * 32), ptx::umma_smem_desc_sw128(w_addr + k * 32), idesc,
                           (kb | k) ? 1u : 0u);
          ptx::umma_commit(&wempty[slot]);
        }
        ptx::umma_commit(&zempty[s]);
        ptx::umma_commit(&tfull[s]);
      }
    }
  } else if (warp >= 10) {
    // ===================== operand builders: z = modulate(LN(A[r] + B[c] + G[pair])) -> swizzled smem =====================
    const int bw = warp - 10;
    int cur_r = -1;
    uint2 araw[2] = {make_uint2(0, 0), make_uint2(0, 0)};       // A[r] kept packed (bf16 x 8)
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      uint8_t* ztile = smZ + s * kZBytes;
      ptx::mbar_wait(&zempty[s], ((it >> 1) & 1) ^ 1);
#pragma unroll 1
      for (int q = 0; q < kRowsPerBuilder; q += kBatch) {
        const int lr0 = bw * kRowsPerBuilder + q;
        const int d0 = tile * TM + lr0;
        if (d0 >= a.Md) break;
        // all loads of the batch are issued before the first use: kBatch rows (1.5 KB) in flight per warp
        int4 info[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) info[j] = __ldg(a.dir_info + min(d0 + j, a.Md - 1));
        uint2 braw[kBatch][2], graw[kBatch][2];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const bf16* br = a.ab + static_cast<size_t>(info[j].z) * 512 + 256;
          const bf16* gr = a.gp + static_cast<size_t>(info[j].x) * 256;
          braw[j][0] = *reinterpret_cast<const uint2*>(br + 4 * lane);
          braw[j][1] = *reinterpret_cast<const uint2*>(br + 128 + 4 * lane);
          graw[j][0] = *reinterpret_cast<const uint2*>(gr + 4 * lane);
          graw[j][1] = *reinterpret_cast<const uint2*>(gr + 128 + 4 * lane);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          if (d0 + j >= a.Md) break;
          if (info[j].y != cur_r) {            // new source atom (rows are source-major): refresh A[r]
            cur_r = info[j].y;
            const bf16* arow = a.ab + static_cast<size_t>(cur_r) * 512;
            araw[0] = *reinterpret_cast<const uint2*>(arow + 4 * lane);
            araw[1] = *reinterpret_cast<const uint2*>(arow + 128 + 4 * lane);
          }
          float v[8], t[8];
          unpack8(araw, v);
          unpack8(braw[j], t);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] += t[k];
          unpack8(graw[j], t);
          float sum = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            v[k] += t[k];
            sum += v[k];
          }
          const float mean = warp_sum(sum) * (1.0f / 256.0f);
          float sq = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            v[k] -= mean;
            sq += v[k] * v[k];
          }
          const float is = rsqrtf(warp_sum(sq) * (1.0f / 256.0f) + 1e-6f);
          // shift / scale of the molecule: L1-resident (consecutive rows share the molecule)
          const float* ar = a.ada + static_cast<size_t>(info[j].w) * ADA_LD + ADA_COORD;
          float sh[8], sc[8];
          load8_f32(ar, lane, sh);
          load8_f32(ar + 256, lane, sc);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = (v[k] * is) * (1.0f + sc[k]) + sh[k];
          store_z_row(ztile, lr0 + j, lane, v);
        }
      }
      ptx::fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&zfull[s]);
    }
  } else {
    // ===================== epilogue: SiLU -> coord_mlp.2 -> tanh -> adjacency-weighted mean =====================
    const int g = warp >> 2, wq = warp & 3;
    for (int it = g; it < my_tiles; it += 2) {
      const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int row = tile * TM + wq * 32 + lane;
      const bool row_ok = row < a.Md;
      uint8_t fl = 0;
      if (row_ok) fl = a.pflags[__ldg(a.dir_info + row).x];       // in flight while the accumulator is produced
      ptx::mbar_wait(&tfull[g], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(g * TK);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < TK; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld32_sync(t_addr + c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 w = swc2[c + i];
          const float u = act_silu_half<true>(__uint_as_float(v[i]) + w.w);   // coord_mlp.0 packed with 0.5 W, 0.5 b
          s0 = fmaf(u, w.x, s0);
          s1 = fmaf(u, w.y, s1);
          s2 = fmaf(u, w.z, s2);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty[g]);
      if (row_ok) {
        const float a2 = (fl & 1) ? 1.f : 0.f, asp = (fl & 2) ? 1.f : 0.f;
        a.wdir[row] = (act_tanh<true>(s0) + act_tanh<true>(s1) * a2 + act_tanh<true>(s2) * asp) / 3.0f;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int coord_fused_launch(DsContext* ctx, const Plan& plan, const void* ab, const void* gp, const float* ada_l,
                       const uint8_t* pflags, const void* wc1, const float* bc1, const float* wc2, float* wdir,
                       cudaStream_t s) {
  const int Md = 2 * plan.Mp;
  if (Md <= 0) return DS_OK;
  DS_CHECK(plan.dir_info != nullptr, DS_ERR_INVALID, "coord_fused: plan has no directed-edge table");
  static bool attr_set[64] = {};            // the attribute is per device: one flag per device ordinal
  if (!attr_set[ctx->device & 63]) {
    DS_CUDA_CHECK(cudaFuncSetAttribute(coord_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set[ctx->device & 63] = true;
  }
  CUtensorMap tmW;
  DS_TRY(ds_make_tmap_2d(ctx, &tmW, wc1, TK, TK, TK, KB, TK, false));
  CoordArgs a;
  a.ab = reinterpret_cast<const bf16*>(ab);
  a.gp = reinterpret_cast<const bf16*>(gp);
  a.ada = ada_l;
  a.pflags = pflags;
  a.dir_info = plan.dir_info;
  a.bias = bc1;
  a.wc2 = wc2;
  a.wdir = wdir;
  a.Md = Md;
  const int tiles = (Md + TM - 1) / TM;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  ds_launch(coord_fused_kernel, dim3(grid), dim3(kThreads), kSmem, s, tmW, a);
  DS_CUDA_CHECK(cudaGetLastError());
  ctx->launch_count++;
  return DS_OK;
}
