// Fused SE(3)-equivariant coordinate head of one EquivariantMixBlock (models/dmt.py:37-60, MultiCondEquiUpdate) as ONE
// kernel on CTA PAIRS (thread-block clusters of 2, tcgen05 cta_group::2):
//
//   per unordered pair p = (i < j):     G = input_lin[:, e | dist] . X[p]                 MMA1: [128 pairs, 128] x [256, 128]^T
//   per directed edge d = (r -> c):     y = A[r] + B[c] + G[p]                            A, B = per-atom parts of input_lin (hoisted)
//                                       z = modulate(LayerNorm(y), csh[mol], csc[mol])    dmt.py:42-44
//                                       u = SiLU(coord_mlp.0 z + b)                       MMA2: [128 edges, 256] x [256, 256]^T
//                                       w = mean(tanh(coord_mlp.2 u) * [1, adj2d, adjsp]) dmt.py:46-51
//
// It replaces three launches of the layer-wise pipeline (the `gp` GEMM, k_coord_ln_async and the COORD GEMM) and their
// HBM round trips (G: 83 MB out + 2 x 83 MB in, Z: 166 MB out + 166 MB in per block at batch 1024): G lives in TMEM, the
// LayerNorm'd operand Z in shared memory, only X (41 MB), the L2-resident per-atom table `ab` and the 4-byte results
// cross HBM.
//
// The CTA pair works on ONE tile of 128 pairs at a time; CTA rank 0 owns the forward edges (i -> j) of those pairs, rank 1
// the reverse edges (j -> i).  Both MMAs are cta_group::2 instructions of M = 256 (128 rows per CTA, rows of the two CTAs
// = the two directions), N = 256, issued by one thread of the leader CTA: each CTA keeps only HALF of the two weight
// matrices resident (coord_mlp.0: 64 KB, input_lin[e | dist]: 32 KB), which is what makes room for the 64 KB operand
// tile, and every weight byte read from shared memory feeds 256 rows.
//
// Per CTA: 10 warps.
//   warps 0-3 / 4-7   compute groups: even / odd tiles, one THREAD per directed edge (= TMEM lane).  A tile goes
//                     G (TMEM) -> pass A: LayerNorm statistics -> pass B: normalise + modulate -> Z (shared, SWIZZLE_128B)
//                     -> [MMA2 overwrites G's columns] -> epilogue: SiLU, 3 dot products, tanh, adjacency mean -> w[d].
//                     While one group runs its epilogue the other builds the next tile (two TMEM stages of 256 columns).
//   warp 8            loader: weights once; per tile the X tile (TMA, counted on the leader's barrier) and the "window":
//                     the lane-varying halves of the `ab` rows of up to 56 consecutive atoms starting at the tile's first
//                     atom (TMA, swizzled, so that 32 lanes reading 32 different atoms do not bank-conflict).
//   warp 9            TMEM owner; in the leader CTA also the MMA issuer for both CTAs.
// Every mbarrier wait is bounded: a protocol error traps with a message instead of hanging the GPU.
#include "kernels.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int TM = 128;                  // pairs per tile = rows per CTA of the M = 256 MMAs
constexpr int kKb = 16 * 1024;           // one k-block tile: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kWinRows = 56;             // atoms in the window
constexpr int kWinBox = kWinRows * 128;  // 7168 B: 56 rows x 64 bf16
constexpr int OFF_WC1 = 0;               // coord_mlp.0 rows [128 rank, +128): 4 k-blocks
constexpr int OFF_WE = 4 * kKb;          // input_lin[e | dist] rows [128 rank, +128): 2 k-blocks
constexpr int OFF_Z = 6 * kKb;           // operand tile of MMA2: 4 k-blocks
constexpr int OFF_X = 10 * kKb;          // operand tile of MMA1: 2 k-blocks
constexpr int OFF_WIN = 12 * kKb;        // 4 boxes of 56 atoms x 64 channels
constexpr int OFF_TAB = OFF_WIN + 4 * kWinBox;   // [256] float4: coord_mlp.2 rows + coord_mlp.0 bias per column pair
constexpr int OFF_BAR = OFF_TAB + 4096;
constexpr int kSmem = OFF_BAR + 256;
constexpr int kThreads = 320;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

struct CoordHeadArgs {
  const bf16* ab;              // [Mn,512]  A = cols 0..255 (with input_lin.bias), B = cols 256..511
  const float* ada;            // adaLN table pre-offset to block + ADA_COORD: shift [256] | scale [256]; row stride ADA_LD
  const uint8_t* pflags;       // [Mp] adjacency bits of the pair (bit 0 adj2d, bit 1 adjsp)
  const uint32_t* pair_info;   // mol << 12 | i << 6 | j
  const int2* pair_rows;       // atom rows of (i, j)
  const int* n_atoms;          // [B]
  const int* poff;             // [B+1]
  const float* bc1;            // coord_mlp.0 bias (halved together with the weight)
  const float* wc2;            // coord_mlp.2 [3,256]
  float* wdir;                 // [2 Mp] per directed edge, source-major
  int Mp;
};

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// bounded wait: tag identifies the barrier in the message
__device__ __forceinline__ void wait_guard(uint64_t* bar, uint32_t parity, int tag, int it) {
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("coord_head_kernel: wait timeout tag=%d it=%d block=%d thread=%d parity=%u\n", tag, it, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg128f(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// y[32] = G (TMEM chunk) + uniform row chunk (global) + lane-varying row chunk (window or global) for channels [c0, c0+32)
__device__ __forceinline__ void load_y32(uint32_t t_addr, int c0, const bf16* urow, const bf16* vrow_g, const uint8_t* win, int wr,
                                         bool in_win, float (&y)[32]) {
  uint4 u[4], v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) u[q] = ldg128(urow + c0 + q * 8);
  if (in_win) {
    const uint8_t* wb = win + (c0 >> 6) * kWinBox + wr * 128;
    const int ch0 = (c0 & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const uint4*>(wb + (((ch0 + q) ^ (wr & 7)) << 4));
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = ldg128(vrow_g + c0 + q * 8);
  }
  uint32_t acc[32];
  ptx::tmem_ld32_sync(t_addr + c0, acc);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float uf[8], vf[8];
    unpack8(u[q], uf);
    unpack8(v[q], vf);
#pragma unroll
    for (int k = 0; k < 8; ++k) y[q * 8 + k] = __uint_as_float(acc[q * 8 + k]) + (uf[k] + vf[k]);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
coord_head_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                  const __grid_constant__ CUtensorMap tmWc1, const __grid_constant__ CUtensorMap tmAB, CoordHeadArgs a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  float4* swc2 = reinterpret_cast<float4*>(smem + OFF_TAB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* wl_full = bars + 0;      // local: this CTA's weight halves have landed (tx)
  uint64_t* w_ready = bars + 1;      // leader: both CTAs' weights are in place (2 arrivals)
  uint64_t* x_full = bars + 2;       // leader: X tiles of both CTAs have landed (1 arrival + tx of both)
  uint64_t* g_full = bars + 3;       // [2] both: MMA1 of the stage is complete (commit multicast)
  uint64_t* z_full = bars + 5;       // leader: both CTAs' operand tiles are written (2 arrivals)
  uint64_t* acc_full = bars + 6;     // [2] both: MMA2 of the stage is complete (commit multicast)
  uint64_t* tmem_free = bars + 8;    // [2] leader: both CTAs have drained the stage (2 arrivals)
  uint64_t* win_full = bars + 10;    // local: window landed (tx)
  uint64_t* win_free = bars + 11;    // local: window consumed (1 arrival)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int ncl = static_cast<int>(gridDim.x >> 1), cid = static_cast<int>(blockIdx.x >> 1);
  const int n_tiles = (a.Mp + TM - 1) / TM;
  const int my_n = (n_tiles - cid + ncl - 1) / ncl;

  if (warp == 8 && lane == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmWe);
    ptx::prefetch_tmap(&tmWc1);
    ptx::prefetch_tmap(&tmAB);
    ptx::mbar_init(wl_full, 1);
    ptx::mbar_init(w_ready, 2);
    ptx::mbar_init(x_full, 1);
    ptx::mbar_init(z_full, 2);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&g_full[i], 1);
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&tmem_free[i], 2);
    }
    ptx::mbar_init(win_full, 1);
    ptx::mbar_init(win_free, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc_2sm<512>(tmem_slot);
  if (threadIdx.x < 256) {     // per column PAIR (2p, 2p+1): [2p] = (w0, w0', w1, w1'), [2p+1] = (w2, w2', bias, bias')
    const int c2 = threadIdx.x & ~1;
    swc2[threadIdx.x] = (threadIdx.x & 1) ? make_float4(a.wc2[512 + c2], a.wc2[513 + c2], a.bc1[c2], a.bc1[c2 + 1])
                                          : make_float4(a.wc2[c2], a.wc2[c2 + 1], a.wc2[256 + c2], a.wc2[257 + c2]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();           // both CTAs' barriers are initialised before anyone arrives remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 8) {
    // ===================== loader =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(wl_full, 6 * kKb);
      for (int kb = 0; kb < 4; ++kb) ptx::tma_load_2d(smem + OFF_WC1 + kb * kKb, &tmWc1, wl_full, kb * 64, static_cast<int>(rank) * 128);
      for (int kb = 0; kb < 2; ++kb) ptx::tma_load_2d(smem + OFF_WE + kb * kKb, &tmWe, wl_full, kb * 64, static_cast<int>(rank) * 128);
      wait_guard(wl_full, 0, 1, 0);
      ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(w_ready), 0));
      const uint32_t x_full_leader = ptx::mapa(ptx::smem_u32(x_full), 0);
      for (int it = 0; it < my_n; ++it) {
        const int p0 = (cid + it * ncl) * TM;
        if (it >= 1) wait_guard(&g_full[(it - 1) & 1], ((it - 1) >> 1) & 1, 2, it);     // MMA1(it-1) has consumed the X tile
        if (rank == 0) ptx::mbar_arrive_expect_tx(x_full, 2 * 2 * kKb);                   // bytes of BOTH CTAs' tiles
        ptx::tma_load_2d_2sm(smem + OFF_X, &tmX, x_full_leader, 0, p0);
        ptx::tma_load_2d_2sm(smem + OFF_X + kKb, &tmX, x_full_leader, 64, p0);
        if (it >= 1) wait_guard(win_free, (it - 1) & 1, 3, it);
        const int a_lo = __ldg(&a.pair_rows[min(p0, a.Mp - 1)].x);
        ptx::mbar_arrive_expect_tx(win_full, 4 * kWinBox);
        for (int b = 0; b < 4; ++b) ptx::tma_load_2d(smem + OFF_WIN + b * kWinBox, &tmAB, win_full, (rank ? 0 : 256) + 64 * b, a_lo);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (leader CTA, one thread, for both CTAs) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(256, 256);
      wait_guard(w_ready, 0, 4, 0);
      ptx::tc_fence_after();
      const uint32_t x_addr = ptx::smem_u32(smem + OFF_X), we_addr = ptx::smem_u32(smem + OFF_WE);
      const uint32_t z_addr = ptx::smem_u32(smem + OFF_Z), wc_addr = ptx::smem_u32(smem + OFF_WC1);
      int next1 = 0, next2 = 0;
      uint32_t idle = 0;
      while (next2 < my_n) {
        bool progressed = false;
        if (next1 < my_n && next1 <= next2 + 1) {
          const int s = next1 & 1;
          const bool free_ok = next1 < 2 || ptx::mbar_try_wait_cluster(&tmem_free[s], ((next1 >> 1) - 1) & 1);
          if (free_ok && ptx::mbar_try_wait_cluster(x_full, next1 & 1)) {
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + static_cast<uint32_t>(s * 256);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma2_bf16(d, ptx::umma_smem_desc_sw128(x_addr + kb * kKb + k * 32), ptx::umma_smem_desc_sw128(we_addr + kb * kKb + k * 32),
                                idesc, (kb | k) ? 1u : 0u);
            ptx::umma2_commit_mc(&g_full[s], 3);
            ++next1;
            progressed = true;
          }
        }
        if (next2 < next1 && ptx::mbar_try_wait_cluster(z_full, next2 & 1)) {
          ptx::tc_fence_after();
          const int s = next2 & 1;
          const uint32_t d = tmem_base + static_cast<uint32_t>(s * 256);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma2_bf16(d, ptx::umma_smem_desc_sw128(z_addr + kb * kKb + k * 32), ptx::umma_smem_desc_sw128(wc_addr + kb * kKb + k * 32),
                              idesc, (kb | k) ? 1u : 0u);
          ptx::umma2_commit_mc(&acc_full[s], 3);
          ++next2;
          progressed = true;
        }
        if (progressed) idle = 0;
        else if (++idle > (1u << 24)) {
          printf("coord_head_kernel: MMA issuer stalled next1=%d next2=%d of %d block=%d\n", next1, next2, my_n, blockIdx.x);
          __trap();
        }
      }
    }
  } else {
    // ===================== compute groups: one thread per directed edge =====================
    const int g = warp >> 2, wq = warp & 3, r = wq * 32 + lane;     // r = row inside the tile = TMEM lane
    const uint32_t z_full_leader = ptx::mapa(ptx::smem_u32(z_full), 0);
    uint8_t* zbuf = smem + OFF_Z;
    const uint8_t* win = smem + OFF_WIN;
    for (int it = g; it < my_n; it += 2) {
      const int s = g;
      const uint32_t sphase = (it >> 1) & 1;
      const int p0 = (cid + it * ncl) * TM;
      const int p = p0 + r;
      const bool ok = p < a.Mp;
      const int pc = ok ? p : a.Mp - 1;
      const int2 rows = __ldg(a.pair_rows + pc);
      const uint32_t info = __ldg(a.pair_info + pc);
      const int mol = info >> 12, ai = (info >> 6) & 63, aj = info & 63;
      const int a_lo = __ldg(&a.pair_rows[min(p0, a.Mp - 1)].x);
      const int nat = __ldg(a.n_atoms + mol), pb = __ldg(a.poff + mol);
      const uint8_t fl = __ldg(a.pflags + pc);
      const float* ar = a.ada + static_cast<size_t>(mol) * ADA_LD;
      // rank 0: edge i -> j: y = A[i] + B[j] + G;   rank 1: edge j -> i: y = A[j] + B[i] + G.  The atom that is (nearly)
      // uniform over the lanes of a warp is i, the lane-varying one j: its half-row comes from the window.
      const bf16* urow = a.ab + static_cast<size_t>(rows.x) * 512 + (rank ? 256 : 0);
      const bf16* vrow = a.ab + static_cast<size_t>(rows.y) * 512 + (rank ? 0 : 256);
      const int wr = rows.y - a_lo;
      const bool in_win = wr >= 0 && wr < kWinRows;
      const size_t d_out = static_cast<size_t>(2) * pb + (rank ? static_cast<size_t>(aj) * (nat - 1) + ai : static_cast<size_t>(ai) * (nat - 1) + (aj - 1));
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(s * 256);

      wait_guard(&g_full[s], sphase, 5, it);
      ptx::tc_fence_after();
      // a parity wait may run at most one phase ahead of the barrier: the window of tile it-1 must have been consumed (hence
      // landed) before win_full can be asked about tile it
      if (it >= 1) wait_guard(win_free, (it - 1) & 1, 9, it);
      wait_guard(win_full, it & 1, 6, it);
      // ---- pass A: LayerNorm statistics of y = G + A + B over the 256 channels (thread-local)
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        float y[32];
        load_y32(t_addr, c0, urow, vrow, win, wr, in_win, y);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          sum += y[k];
          sq = fmaf(y[k], y[k], sq);
        }
      }
      const float mean = sum * (1.0f / 256.0f);
      const float is = rsqrtf(fmaxf(sq * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-6f);
      const float nm = -mean * is;
      // the operand buffer is free once MMA2 of the previous tile (the other group's stage) has completed
      if (it >= 1) wait_guard(&acc_full[s ^ 1], ((it - 1) >> 1) & 1, 7, it);
      // ---- pass B: normalise, modulate, bf16, SWIZZLE_128B K-major operand row
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        float y[32];
        load_y32(t_addr, c0, urow, vrow, win, wr, in_win, y);
        uint8_t* zrow = zbuf + (c0 >> 6) * kKb + r * 128;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 sh0 = ldg128f(ar + c0 + q * 8), sh1 = ldg128f(ar + c0 + q * 8 + 4);
          const float4 sc0 = ldg128f(ar + 256 + c0 + q * 8), sc1 = ldg128f(ar + 256 + c0 + q * 8 + 4);
          const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
          const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
          float z[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) z[k] = fmaf(fmaf(y[q * 8 + k], is, nm), 1.0f + scv[k], shv[k]);
          *reinterpret_cast<uint4*>(zrow + (((ch0 + q) ^ (r & 7)) << 4)) =
              make_uint4(pack2(z[0], z[1]), pack2(z[2], z[3]), pack2(z[4], z[5]), pack2(z[6], z[7]));
        }
      }
      ptx::tc_fence_before();            // this thread's reads of G precede the MMA2 that overwrites those columns
      ptx::fence_proxy_async_smem();     // operand row visible to the tensor core
      group_sync(g);
      if (r == 0) {
        ptx::mbar_arrive(win_free);
        ptx::mbar_arrive_cluster(z_full_leader);
      }
      // ---- epilogue: u = SiLU(2 (acc + b/2)) ; s_o = wc2[o] . u ; w = mean(tanh(s) * [1, adj2d, adjsp])
      wait_guard(&acc_full[s], sphase, 8, it);
      ptx::tc_fence_after();
      float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0;
#pragma unroll 1
      for (int c = 0; c < 256; c += 32) {
        uint32_t acc[32];
        ptx::tmem_ld32_sync(t_addr + c, acc);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float4 wa = swc2[c + i], wb = swc2[c + i + 1];
          const float2 h = fadd2(make_float2(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1])), make_float2(wb.z, wb.w));
          const float2 v = ffma2(h, make_float2(act_tanh<true>(h.x), act_tanh<true>(h.y)), h);   // SiLU(2h) = h + h tanh(h)
          q0 = ffma2(v, make_float2(wa.x, wa.y), q0);
          q1 = ffma2(v, make_float2(wa.z, wa.w), q1);
          q2 = ffma2(v, make_float2(wb.x, wb.y), q2);
        }
      }
      const float s0 = q0.x + q0.y, s1 = q1.x + q1.y, s2 = q2.x + q2.y;
      if (ok) {
        const float a2 = (fl & 1) ? 1.f : 0.f, asp = (fl & 2) ? 1.f : 0.f;
        a.wdir[d_out] = (act_tanh<true>(s0) + act_tanh<true>(s1) * a2 + act_tanh<true>(s2) * asp) / 3.0f;
      }
      ptx::tc_fence_before();
      group_sync(g);
      if (r == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tmem_free[s]), 0));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();           // the peer may still be read by / written from the leader's MMAs until here
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ----------------------------------------------------------------------------- probe: one cta_group::2 MMA tile
// out[256, 256] (fp32) = A[256, K] . W[256, K]^T with ONE cluster: CTA r holds rows [128 r, +128) of A and of W.  Pins the
// operand / accumulator split of the CTA-pair MMA that coord_head_kernel relies on (tests/test_gemm_gpu.py).
constexpr int kProbeMaxKb = 4;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, float* out, int kbs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smA = smem;
  uint8_t* smW = smem + kProbeMaxKb * kKb;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kProbeMaxKb * kKb);
  uint64_t* ab_full = bars;       // leader: operands of both CTAs landed
  uint64_t* d_full = bars + 1;    // both: MMA complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();
    ptx::mbar_init(ab_full, 1);
    ptx::mbar_init(d_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_2sm<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t leader_bar = ptx::mapa(ptx::smem_u32(ab_full), 0);
    if (rank == 0) ptx::mbar_arrive_expect_tx(ab_full, 2 * 2 * kbs * kKb);
    for (int kb = 0; kb < kbs; ++kb) {
      ptx::tma_load_2d_2sm(smA + kb * kKb, &tmA, leader_bar, kb * 64, static_cast<int>(rank) * 128);
      ptx::tma_load_2d_2sm(smW + kb * kKb, &tmW, leader_bar, kb * 64, static_cast<int>(rank) * 128);
    }
    if (rank == 0) {
      wait_guard(ab_full, 0, 20, 0);
      ptx::tc_fence_after();
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(256, 256);
      for (int kb = 0; kb < kbs; ++kb)
        for (int k = 0; k < 4; ++k)
          ptx::umma2_bf16(tmem_base, ptx::umma_smem_desc_sw128(ptx::smem_u32(smA + kb * kKb) + k * 32),
                          ptx::umma_smem_desc_sw128(ptx::smem_u32(smW + kb * kKb) + k * 32), idesc, (kb | k) ? 1u : 0u);
      ptx::umma2_commit_mc(d_full, 3);
    }
  }
  __syncwarp();
  wait_guard(d_full, 0, 21, 0);
  ptx::tc_fence_after();
  const int row = static_cast<int>(rank) * 128 + warp * 32 + lane;
  const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < 256; c += 32) {
    uint32_t acc[32];
    ptx::tmem_ld32_sync(t_addr + c, acc);
    for (int i = 0; i < 32; ++i) out[static_cast<size_t>(row) * 256 + c + i] = __uint_as_float(acc[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm<256>(tmem_base);
  }
}

}  // namespace

int coord_head_launch(DsContext* ctx, const Plan& plan, const void* X, const void* ab, const float* ada_l, const uint8_t* pflags,
                      const void* we, const void* wc1, const float* bc1, const float* wc2, float* wdir, cudaStream_t s) {
  if (plan.Mp <= 0) return DS_OK;
  static bool attr_set[64] = {};            // the attribute is per device: one flag per device ordinal
  if (!attr_set[ctx->device & 63]) {
    DS_CUDA_CHECK(cudaFuncSetAttribute(coord_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set[ctx->device & 63] = true;
  }
  CUtensorMap tmX, tmWe, tmWc1, tmAB;
  DS_TRY(ds_make_tmap_2d(ctx, &tmX, X, plan.Mp, 128, 128, 64, TM, false));
  DS_TRY(ds_make_tmap_2d(ctx, &tmWe, we, 256, 128, 128, 64, 128, false));
  DS_TRY(ds_make_tmap_2d(ctx, &tmWc1, wc1, 256, 256, 256, 64, 128, false));
  DS_TRY(ds_make_tmap_2d(ctx, &tmAB, ab, plan.Mn, 512, 512, 64, kWinRows, false));
  CoordHeadArgs a;
  a.ab = reinterpret_cast<const bf16*>(ab);
  a.ada = ada_l + ADA_COORD;
  a.pflags = pflags;
  a.pair_info = plan.pair_info;
  a.pair_rows = plan.pair_rows;
  a.n_atoms = plan.n_atoms;
  a.poff = plan.poff;
  a.bc1 = bc1;
  a.wc2 = wc2;
  a.wdir = wdir;
  a.Mp = plan.Mp;
  const int tiles = (plan.Mp + TM - 1) / TM;
  const int max_cl = ctx->num_sms / 2;
  const int ncl = tiles < max_cl ? tiles : max_cl;
  ds_launch(coord_head_kernel, dim3(2 * ncl), dim3(kThreads), kSmem, s, tmX, tmWe, tmWc1, tmAB, a);
  DS_CUDA_CHECK(cudaGetLastError());
  ctx->launch_count++;
  return DS_OK;
}

int umma2_probe_launch(DsContext* ctx, const void* A, const void* W, float* out, int K, cudaStream_t s) {
  DS_CHECK(K % 64 == 0 && K >= 64 && K <= 64 * kProbeMaxKb, DS_ERR_INVALID, "umma2 probe: K must be 64..%d in steps of 64", 64 * kProbeMaxKb);
  constexpr int smem_bytes = 2 * kProbeMaxKb * kKb + 64;
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    DS_CUDA_CHECK(cudaFuncSetAttribute(umma2_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set[ctx->device & 63] = true;
  }
  CUtensorMap tmA, tmW;
  DS_TRY(ds_make_tmap_2d(ctx, &tmA, A, 256, K, K, 64, 128, false));
  DS_TRY(ds_make_tmap_2d(ctx, &tmW, W, 256, K, K, 64, 128, false));
  ds_launch(umma2_probe_kernel, dim3(2), dim3(128), smem_bytes, s, tmA, tmW, out, K / 64);
  DS_CUDA_CHECK(cudaGetLastError());
  ctx->launch_count++;
  return DS_OK;
}
