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
// The CTA pair works on tiles of 128 pairs; CTA rank 0 owns the forward edges (i -> j) of those pairs, rank 1 the reverse
// edges (j -> i).  All MMAs are cta_group::2 instructions of M = 256 (128 rows per CTA, rows of the two CTAs = the two
// directions) issued by one thread of the leader CTA: each CTA keeps only HALF of the two weight matrices resident
// (coord_mlp.0: 64 KB, input_lin[e | dist]: 32 KB), which is what makes room for the 64 KB operand tile, and every weight
// byte read from shared memory feeds 256 rows.
//
// TMEM (512 columns per CTA): G of the current tile in [0, 256) (MMA1, N = 256); the accumulator of coord_mlp.0 as two column
// HALVES [256, 384) and [384, 512) (two N = 128 MMA2s per tile), drained one after the other.  G(t+1) only waits for pass A(t)
// of both CTAs (g_free), the halves of MMA2(t) for the epilogue of the same half of tile t-1 (acc_free), the operand tile Z
// for MMA2(t-1) (acc_full[1]): build and epilogue run one tile apart without meeting in the TMEM.
//
// Per CTA: 20 warps (640 threads; the register file is re-divided with setmaxnreg: 144 / 72 / 40 registers per thread).
//   warps 0-7    build: TWO threads per directed edge (= TMEM lane; warps w and w + 4 share a lane quadrant and split the 256
//                channels).  Pass A(t): y = G + A + B, LayerNorm statistics, y kept in REGISTERS as packed bf16; pass B(t):
//                normalise + modulate -> operand tile Z in shared memory (SWIZZLE_128B) -> MMA2(t).
//   warps 8-15   epilogue: two threads per directed edge, 64 columns of each accumulator half: SiLU, the three dot products
//                of coord_mlp.2, tanh, adjacency mean -> w[d].  MUFU.TANH-bound; runs next to the build of tile t+1.
//   warp 16      loader: weights once; per tile the X tile (TMA, counted on the leader's barrier) and the "window": the
//                lane-varying halves of the `ab` rows of up to 48 consecutive atoms starting at the tile's first atom (TMA,
//                swizzled, so that 32 lanes reading 32 different atoms do not bank-conflict); atoms beyond the window (runs
//                of tiny molecules) are read from global memory.
//   warp 17      TMEM owner; in the leader CTA also the MMA issuer for both CTAs.
//   warps 18-19  idle: their registers go to the build warps.
// Every mbarrier wait is bounded: a protocol error traps with a message instead of hanging the GPU.
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "ptx_sm100.cuh"

namespace {

constexpr int TM = 128;                  // pairs per tile = rows per CTA of the M = 256 MMAs
constexpr int kKb = 16 * 1024;           // one k-block tile: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kWinRows = 48;             // atoms in the window
constexpr int kWinBox = kWinRows * 128;  // 6144 B: 48 rows x 64 bf16
constexpr int OFF_WC1 = 0;               // coord_mlp.0 rows [128 rank, +128): 4 k-blocks
constexpr int OFF_WE = 4 * kKb;          // input_lin[e | dist] rows [128 rank, +128): 2 k-blocks
constexpr int OFF_Z = 6 * kKb;           // operand tile of MMA2: 4 k-blocks
constexpr int OFF_X = 10 * kKb;          // operand tile of MMA1: 2 k-blocks
constexpr int OFF_WIN = 12 * kKb;        // 4 boxes of 48 atoms x 64 channels
constexpr int OFF_TAB = OFF_WIN + 4 * kWinBox;   // [256] float4: per column PAIR (2p, 2p+1): [2p] = (w0, w0', w1, w1'), [2p+1] = (w2, w2', bias, bias')
constexpr int OFF_STAT = OFF_TAB + 4096;         // [2][128] float2 LayerNorm partials of the two channel halves
constexpr int OFF_PART = OFF_STAT + 2048;        // [128] float4 epilogue partial sums of the upper channel half
constexpr int OFF_BAR = OFF_PART + 2048;
constexpr int kSmem = OFF_BAR + 256;
constexpr int kThreads = 640;             // 8 build warps, 8 epilogue warps, loader, MMA issuer, 2 idle (register donors)
static_assert(kSmem <= 227 * 1024, "shared memory budget");

struct CoordHeadArgs {
  const bf16* ab;              // [Mn,512]  A = cols 0..255 (with input_lin.bias), B = cols 256..511
  const bf16* cmod;            // [B][512] modulate vectors of this block as bf16: shift [256] | 1 + scale [256] (k_coord_mod)
  const uint8_t* pflags;       // [Mp] adjacency bits of the pair (bit 0 adj2d, bit 1 adjsp)
  const uint32_t* pair_info;   // mol << 12 | i << 6 | j
  const int2* pair_rows;       // atom rows of (i, j)
  const int* n_atoms;          // [B]
  const int* poff;             // [B+1]
  const float* bc1;            // coord_mlp.0 bias (halved together with the weight)
  const float* wc2;            // coord_mlp.2 [3,256]
  float* wdir;                 // [2 Mp] per directed edge, source-major
  int Mp;
  long long* dbg;              // optional [8 x tiles of block 0] clock stamps of thread 0 (DS_COORD_DBG), else null
};

// bounded wait: tag identifies the barrier in the message.  CTA-scope acquire on purpose: a cluster-scope acquire compiles to
// TRYWAIT + CCTL.IVALL (the whole L1 is invalidated at every successful wait; measured: 30 % of the kernel's stall samples and
// every per-atom row re-fetched from L2).  What the waits protect is TMEM, TMA-written shared memory or operand tiles read
// by the tensor core, none of which lives in L1; the remote arrivals carry release.cluster.
__device__ __forceinline__ void wait_guard(uint64_t* bar, uint32_t parity, int tag, int it) {
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    __nanosleep(40);            // a spinning warp takes issue slots from the warps it is waiting for
    if (++spins > (1u << 22)) {
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
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
coord_head_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                  const __grid_constant__ CUtensorMap tmWc1, const __grid_constant__ CUtensorMap tmAB, CoordHeadArgs a) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* wl_full = bars + 0;      // local: this CTA's weight halves have landed (tx)
  uint64_t* w_ready = bars + 1;      // leader: both CTAs' weights are in place (2 arrivals)
  uint64_t* x_full = bars + 2;       // leader: X tiles of both CTAs have landed (1 arrival + tx of both)
  uint64_t* g_full = bars + 3;       // both: MMA1 of the tile is complete (commit multicast)
  uint64_t* g_free = bars + 4;       // leader: every build warp of both CTAs has read G out of the TMEM (16 arrivals)
  uint64_t* z_full = bars + 5;       // leader: both CTAs' operand tiles are written (2 arrivals)
  uint64_t* acc_full = bars + 6;     // [2] both: MMA2 of the tile's column half is complete (commit multicast)
  uint64_t* acc_free = bars + 8;     // [2] leader: every epilogue warp of both CTAs has drained the column half (16 arrivals)
  uint64_t* win_full = bars + 10;    // local: window landed (tx)
  uint64_t* win_free = bars + 11;    // local: window consumed (1 arrival)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int ncl = static_cast<int>(gridDim.x >> 1), cid = static_cast<int>(blockIdx.x >> 1);
  const int n_tiles = (a.Mp + TM - 1) / TM;
  const int my_n = (n_tiles - cid + ncl - 1) / ncl;

  if (warp == 16 && lane == 0) {
    if (ptx::smem_u32(smem) & 1023u) __trap();
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmWe);
    ptx::prefetch_tmap(&tmWc1);
    ptx::prefetch_tmap(&tmAB);
    ptx::mbar_init(wl_full, 1);
    ptx::mbar_init(w_ready, 2);
    ptx::mbar_init(x_full, 1);
    ptx::mbar_init(z_full, 2);
    ptx::mbar_init(g_full, 1);
    ptx::mbar_init(g_free, 16);                // one arrival per build warp of both CTAs
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_free[i], 16);        // one arrival per epilogue warp of both CTAs
    }
    ptx::mbar_init(win_full, 1);
    ptx::mbar_init(win_free, 8);             // one arrival per build warp
    ptx::fence_barrier_init();
  }
  if (warp == 17) ptx::tmem_alloc_2sm<512>(tmem_slot);
  if (threadIdx.x < 256) {     // epilogue table (fp32: unpacking a bf16 table costs more issue slots than the smaller loads save)
    const int c2 = threadIdx.x & ~1;
    reinterpret_cast<float4*>(smem + OFF_TAB)[threadIdx.x] =
        (threadIdx.x & 1) ? make_float4(a.wc2[512 + c2], a.wc2[513 + c2], a.bc1[c2], a.bc1[c2 + 1])
                          : make_float4(a.wc2[c2], a.wc2[c2 + 1], a.wc2[256 + c2], a.wc2[257 + c2]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();           // both CTAs' barriers are initialised before anyone arrives remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 16) {
    // ===================== loader =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(wl_full, 6 * kKb);
      for (int kb = 0; kb < 4; ++kb) ptx::tma_load_2d(smem + OFF_WC1 + kb * kKb, &tmWc1, wl_full, kb * 64, static_cast<int>(rank) * 128);
      for (int kb = 0; kb < 2; ++kb) ptx::tma_load_2d(smem + OFF_WE + kb * kKb, &tmWe, wl_full, kb * 64, static_cast<int>(rank) * 128);
      wait_guard(wl_full, 0, 1, 0);
      ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(w_ready), 0));
      const uint32_t x_full_leader = ptx::mapa(ptx::smem_u32(x_full), 0);
      for (int it = 0; it < my_n; ++it) {
        const int p0 = (cid + it * ncl) * TM;
        if (it >= 1) wait_guard(g_full, (it - 1) & 1, 2, it);                          // MMA1(it-1) has consumed the X tile
        if (rank == 0) ptx::mbar_arrive_expect_tx(x_full, 2 * 2 * kKb);                   // bytes of BOTH CTAs' tiles
        ptx::tma_load_2d_2sm(smem + OFF_X, &tmX, x_full_leader, 0, p0);
        ptx::tma_load_2d_2sm(smem + OFF_X + kKb, &tmX, x_full_leader, 64, p0);
        if (it >= 1) wait_guard(win_free, (it - 1) & 1, 3, it);
        const int a_lo = __ldg(&a.pair_rows[min(p0, a.Mp - 1)].x);
        ptx::mbar_arrive_expect_tx(win_full, 4 * kWinBox);
        for (int b = 0; b < 4; ++b) ptx::tma_load_2d(smem + OFF_WIN + b * kWinBox, &tmAB, win_full, (rank ? 0 : 256) + 64 * b, a_lo);
      }
    }
  } else if (warp == 17) {
    // ===================== MMA issuer (leader CTA, one thread, for both CTAs) =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (rank == 0 && lane == 0) {
      // TMEM: G in columns [0, 256); the accumulator of coord_mlp.0 in two column HALVES [256, 384) and [384, 512), produced by
      // two N = 128 MMAs per tile and drained one after the other.  G(t+1) therefore only waits for pass A(t) (g_free), not
      // for the epilogue of tile t-1 as it did when G and the accumulator shared two 256-column stages (the build warps spent
      // 30 % of their samples waiting for G, the epilogue warps 41 % waiting for the accumulator: profiles/r2_ncu_coord_head.md)
      constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(256, 256), idesc2 = ptx::umma_idesc_bf16(256, 128);
      wait_guard(w_ready, 0, 4, 0);
      ptx::tc_fence_after();
      const uint32_t x_addr = ptx::smem_u32(smem + OFF_X), we_addr = ptx::smem_u32(smem + OFF_WE);
      const uint32_t z_addr = ptx::smem_u32(smem + OFF_Z), wc_addr = ptx::smem_u32(smem + OFF_WC1);
      int next1 = 0, next2 = 0;          // next2 counts column halves: tile = next2 >> 1, half = next2 & 1
      uint32_t idle = 0;
      while (next2 < 2 * my_n) {
        bool progressed = false;
        if (next1 < my_n) {
          const bool free_ok = next1 < 1 || ptx::mbar_try_wait(g_free, (next1 - 1) & 1);
          if (free_ok && ptx::mbar_try_wait(x_full, next1 & 1)) {
            ptx::tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma2_bf16(tmem_base, ptx::umma_smem_desc_sw128(x_addr + kb * kKb + k * 32), ptx::umma_smem_desc_sw128(we_addr + kb * kKb + k * 32),
                                idesc1, (kb | k) ? 1u : 0u);
            ptx::umma2_commit_mc(g_full, 3);
            ++next1;
            progressed = true;
          }
        }
        {
          const int t2 = next2 >> 1, h = next2 & 1;
          const bool free_ok = t2 < 1 || ptx::mbar_try_wait(&acc_free[h], (t2 - 1) & 1);
          if (t2 < next1 && free_ok && ptx::mbar_try_wait(z_full, t2 & 1)) {
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + 256u + static_cast<uint32_t>(h * 128);
            // column half h = rows [64 h, +64) of each CTA's half of coord_mlp.0: accumulator column c < 64 is output channel
            // 64 h + c (leader's rows), column 64 + c is channel 128 + 64 h + c (peer's rows)
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma2_bf16(d, ptx::umma_smem_desc_sw128(z_addr + kb * kKb + k * 32),
                                ptx::umma_smem_desc_sw128(wc_addr + kb * kKb + h * (64 * 128) + k * 32), idesc2, (kb | k) ? 1u : 0u);
            ptx::umma2_commit_mc(&acc_full[h], 3);
            ++next2;
            progressed = true;
          }
        }
        if (progressed) idle = 0;
        else if (++idle > (1u << 24)) {
          printf("coord_head_kernel: MMA issuer stalled next1=%d next2=%d of %d block=%d\n", next1, next2, my_n, blockIdx.x);
          __trap();
        }
      }
    }
  } else if (warp < 8) {
    // ===================== build warps: TWO threads per directed edge (128 channels each) =====================
    //   pass A(t)   y = G + A + B for the thread's 128 channels -> LayerNorm partial statistics; y stays in REGISTERS as packed
    //               bf16 (64 registers), so neither G, nor the window, nor the per-atom rows are needed again
    //   pass B(t)   normalise + modulate from the registers -> operand tile Z -> MMA2(t)
    // and straight on to tile t+1 while the epilogue warps drain tile t: the build is bound by TMEM reads, the load/store
    // return path and issue slots, the epilogue by MUFU.TANH, so the two overlap on one SM.
    // register pool of the CTA = 640 threads x 96: 256 x 144 (build) + 256 x 72 (epilogue) + 128 x 40 (loader, MMA, idle) = 60 416 <= 61 440
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    const int hf = warp >> 2, wq = warp & 3, r = wq * 32 + lane;     // r = row inside the tile = TMEM lane; hf = channel half
    const int cb = hf * 128;
    const uint32_t z_full_leader = ptx::mapa(ptx::smem_u32(z_full), 0), g_free_leader = ptx::mapa(ptx::smem_u32(g_free), 0);
    float2* sstat = reinterpret_cast<float2*>(smem + OFF_STAT);      // [2][128] (sum, sum of squares) of each channel half
    uint8_t* zbuf = smem + OFF_Z;
    const uint8_t* win = smem + OFF_WIN;
    const bool stamp = a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define DS_STAMP(k) do { if (stamp) a.dbg[it * 8 + (k)] = clock64(); } while (0)
    // indices of the thread's edge one tile AHEAD: every dependent global load of a tile (per-atom rows, modulate vectors) then
    // starts from registers, and its L1 prefetch is issued a whole pass earlier
    auto load_idx = [&](int t, int2& rows_o, int& mol_o, int& alo_o) {
      const int q0 = (cid + t * ncl) * TM;
      const int pc = min(q0 + r, a.Mp - 1);
      rows_o = __ldg(a.pair_rows + pc);
      mol_o = __ldg(a.pair_info + pc) >> 12;
      alo_o = __ldg(&a.pair_rows[min(q0, a.Mp - 1)].x);
    };
    int2 rows_n = make_int2(0, 0);
    int mol_n = 0, alo_n = 0;
    if (my_n > 0) load_idx(0, rows_n, mol_n, alo_n);
    for (int it = 0; it < my_n; ++it) {
      DS_STAMP(0);
      uint32_t yreg[64];          // the thread's 128 channels of y, packed bf16 pairs
      const int2 rows = rows_n;
      const int mol = mol_n, a_lo = alo_n;
      if (it + 1 < my_n) load_idx(it + 1, rows_n, mol_n, alo_n);
      const bf16* cm = a.cmod + static_cast<size_t>(mol) * 512 + cb;
      // rank 0: edge i -> j: y = A[i] + B[j] + G;   rank 1: edge j -> i: y = A[j] + B[i] + G.  The atom that is (nearly)
      // uniform over the lanes of a warp is i, the lane-varying one j: its half-row comes from the window.
      const bf16* urow = a.ab + static_cast<size_t>(rows.x) * 512 + (rank ? 256 : 0) + cb;
      const bf16* vrow = a.ab + static_cast<size_t>(rows.y) * 512 + (rank ? 0 : 256) + cb;
      const int wr = rows.y - a_lo;
      const bool in_win = wr >= 0 && wr < kWinRows;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(cb);
      if (it == 0) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(urow));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(urow + 64));
      }
      // the modulate vectors of this thread's channels are wanted by pass B, a whole pass A from here
      asm volatile("prefetch.global.L1 [%0];" ::"l"(cm));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(cm + 64));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(cm + 256));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(cm + 320));
      wait_guard(g_full, it & 1, 5, it);
      ptx::tc_fence_after();
      DS_STAMP(1);
      wait_guard(win_full, it & 1, 6, it);
      DS_STAMP(2);
      // ---- pass A: chunks of 8 channels; the TMEM load of chunk c+1 is in flight while chunk c is consumed.  (Chunks of 16
      // spilled ~22 registers of y to local memory, whose L1 traffic equalled all the global loads of the kernel.)
      float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
      uint32_t accb[2][8];
      ptx::tmem_ld8_nowait(t_addr, accb[0]);
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 8) {
        uint32_t (&acc)[8] = accb[(c0 >> 3) & 1];
        const uint4 u = ldg128(urow + c0);
        uint4 v;
        if (in_win) {
          const uint8_t* wb = win + ((cb + c0) >> 6) * kWinBox + wr * 128;
          const int ch0 = ((cb + c0) & 63) >> 3;
          v = *reinterpret_cast<const uint4*>(wb + ((ch0 ^ (wr & 7)) << 4));
        } else {
          v = ldg128(vrow + c0);
        }
        ptx::tmem_wait_ld8(acc);
        if (c0 + 8 < 128) ptx::tmem_ld8_nowait(t_addr + c0 + 8, accb[((c0 >> 3) + 1) & 1]);
        // A + B as packed bf16 adds (both operands are bf16 already), then fp32: + G, statistics
        const __nv_bfloat162* u2 = reinterpret_cast<const __nv_bfloat162*>(&u);
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 ab2 = __hadd2(u2[k], v2[k]);
          const uint32_t w = *reinterpret_cast<const uint32_t*>(&ab2);
          const float2 y2 = fadd2(make_float2(__uint_as_float(acc[2 * k]), __uint_as_float(acc[2 * k + 1])),
                                  make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)));
          // the statistics are those of the ROUNDED values (the ones pass B normalises); this also ties the pack to pass A:
          // left free, ptxas sinks the 64 packs below the LayerNorm barrier and keeps y as fp32 pairs, spilling ~20 of them
          const uint32_t yb = pack2(y2.x, y2.y);
          const float2 yr = make_float2(__uint_as_float(yb << 16), __uint_as_float(yb & 0xffff0000u));
          sum2 = fadd2(sum2, yr);
          sq2 = ffma2(yr, yr, sq2);
          yreg[(c0 >> 1) + k] = yb;
        }
      }
      sstat[hf * 128 + r] = make_float2(sum2.x + sum2.y, sq2.x + sq2.y);
      if (it + 1 < my_n) {               // the next tile's per-atom half-row (its indices arrived during pass A)
        const bf16* un = a.ab + static_cast<size_t>(rows_n.x) * 512 + (rank ? 256 : 0) + cb;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(un));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(un + 64));
      }
      ptx::tc_fence_before();            // this thread's reads of G precede the MMA1 of the next tile, which overwrites those columns
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive_remote(g_free_leader);                  // this warp has G in registers
        ptx::mbar_arrive(win_free);                              // and is done with the window
      }
      // the two threads of a row sit in warps w and w ^ 4: a 64-thread named barrier per warp pair, not one over all 8 warps
      asm volatile("bar.sync %0, 64;" ::"r"(4 + wq) : "memory");
      const float2 mine = sstat[hf * 128 + r], other = sstat[(hf ^ 1) * 128 + r];
      const float mean = (mine.x + other.x) * (1.0f / 256.0f);
      const float is = rsqrtf(fmaxf((mine.y + other.y) * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-6f);
      DS_STAMP(3);
      // the operand buffer is free once MMA2 of the previous tile has completed
      if (it >= 1) wait_guard(&acc_full[1], (it - 1) & 1, 7, it);      // the second column half is issued last
      DS_STAMP(4);
      // ---- pass B: normalise + modulate from the registers -> bf16 -> SWIZZLE_128B K-major operand row
      // z = (y rstd - mean rstd) (1 + scale) + shift as TWO packed bf16 fmas per channel pair (fp32 inside each, one rounding
      // each): y is bf16 already, the modulate vectors arrive as bf16 pairs, so nothing is unpacked or re-packed.  The
      // per-row factors rstd and -mean rstd are rounded to bf16 (a 2^-9 relative error of the row's scale).
      const __nv_bfloat162 is2 = __float2bfloat162_rn(is), nm2 = __float2bfloat162_rn(-mean * is);
      // chunks of 16 channels with the loads of the next two chunks in flight (a batch of 64 channels at once forced ~20
      // registers of y into local memory)
      uint4 shq[3][2], scq[3][2];
#pragma unroll
      for (int pre = 0; pre < 2; ++pre) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          shq[pre][q] = ldg128(cm + pre * 16 + q * 8);
          scq[pre][q] = ldg128(cm + 256 + pre * 16 + q * 8);
        }
      }
#pragma unroll
      for (int ck = 0; ck < 8; ++ck) {
        if (ck + 2 < 8) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            shq[(ck + 2) % 3][q] = ldg128(cm + (ck + 2) * 16 + q * 8);
            scq[(ck + 2) % 3][q] = ldg128(cm + 256 + (ck + 2) * 16 + q * 8);
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const __nv_bfloat162* sh2 = reinterpret_cast<const __nv_bfloat162*>(&shq[ck % 3][q]);
          const __nv_bfloat162* sc2 = reinterpret_cast<const __nv_bfloat162*>(&scq[ck % 3][q]);
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t w = yreg[ck * 8 + q * 4 + k];
            const __nv_bfloat162 n2 = __hfma2(*reinterpret_cast<const __nv_bfloat162*>(&w), is2, nm2);
            const __nv_bfloat162 z2 = __hfma2(n2, sc2[k], sh2[k]);
            o[k] = *reinterpret_cast<const uint32_t*>(&z2);
          }
          const int cq = (ck & 3) * 2 + q;      // 16-byte chunk of the 64-channel k-block
          *reinterpret_cast<uint4*>(zbuf + (2 * hf + (ck >> 2)) * kKb + r * 128 + ((cq ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      ptx::fence_proxy_async_smem();     // operand row visible to the tensor core
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) ptx::mbar_arrive_cluster(z_full_leader);
      DS_STAMP(5);
    }
  } else if (warp < 16) {
    // ===================== epilogue warps: TWO threads per directed edge, 128 columns each =====================
    //   u = SiLU(2 (acc + b/2)) ; s_o = wc2[o] . u ; w = mean(tanh(s) * [1, adj2d, adjsp])
    // MUFU.TANH-bound (one tanh per accumulator element, measured ~8 per clock per SM): two warps per scheduler keep the
    // special-function unit fed while the build warps use the other pipes.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int hf = (warp - 8) >> 2, wq = warp & 3, r = wq * 32 + lane;
    const float4* stab0 = reinterpret_cast<const float4*>(smem + OFF_TAB);
    float4* spart = reinterpret_cast<float4*>(smem + OFF_PART);      // [128] partial sums of the second thread of a row
    const uint32_t acc_free_leader = ptx::mapa(ptx::smem_u32(acc_free), 0);
    for (int it = 0; it < my_n; ++it) {
      const int p0 = (cid + it * ncl) * TM;
      const int p = p0 + r;
      const bool ok = p < a.Mp;
      const int pc = ok ? p : a.Mp - 1;
      const uint32_t info = __ldg(a.pair_info + pc);
      const int mol = info >> 12, ai = (info >> 6) & 63, aj = info & 63;
      const int nat = __ldg(a.n_atoms + mol), pb = __ldg(a.poff + mol);
      const uint8_t fl = __ldg(a.pflags + pc);
      const size_t d_out = static_cast<size_t>(2) * pb + (rank ? static_cast<size_t>(aj) * (nat - 1) + ai : static_cast<size_t>(ai) * (nat - 1) + (aj - 1));
      float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0, q3 = q0, q4 = q0, q5 = q0;     // two chains per output
      // the accumulator arrives as two column halves; of each, this thread takes columns [64 hf, +64) = output channels
      // 128 hf + 64 ch + [0, 64) (the MMA issuer's comment has the column <-> channel map of an N = 128 pair MMA)
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(256 + ch * 128 + hf * 64);
        const float4* stab = stab0 + hf * 128 + ch * 64;
        wait_guard(&acc_full[ch], it & 1, 8, it);
        ptx::tc_fence_after();
        uint32_t eacc[2][16];
        ptx::tmem_ld16_nowait(t_acc, eacc[0]);
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          uint32_t (&acc)[16] = eacc[(c >> 4) & 1];
          ptx::tmem_wait_ld16(acc);
          if (c + 16 < 64) ptx::tmem_ld16_nowait(t_acc + c + 16, eacc[((c >> 4) + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float4 wa = stab[c + i], wb = stab[c + i + 1];
            const float2 h = fadd2(make_float2(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1])), make_float2(wb.z, wb.w));
            const float2 v = ffma2(h, make_float2(act_tanh<true>(h.x), act_tanh<true>(h.y)), h);   // SiLU(2h) = h + h tanh(h)
            if (i & 2) {
              q3 = ffma2(v, make_float2(wa.x, wa.y), q3);
              q4 = ffma2(v, make_float2(wa.z, wa.w), q4);
              q5 = ffma2(v, make_float2(wb.x, wb.y), q5);
            } else {
              q0 = ffma2(v, make_float2(wa.x, wa.y), q0);
              q1 = ffma2(v, make_float2(wa.z, wa.w), q1);
              q2 = ffma2(v, make_float2(wb.x, wb.y), q2);
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_remote(acc_free_leader + ch * 8);        // this warp's columns of the half are drained
      }
      float s0 = (q0.x + q0.y) + (q3.x + q3.y), s1 = (q1.x + q1.y) + (q4.x + q4.y), s2 = (q2.x + q2.y) + (q5.x + q5.y);
      if (hf == 1) spart[r] = make_float4(s0, s1, s2, 0.f);
      asm volatile("bar.sync %0, 64;" ::"r"(8 + wq) : "memory");       // warps w and w + 4 hold the two halves of a row
      if (hf == 0 && ok) {
        const float4 o = spart[r];
        s0 += o.x; s1 += o.y; s2 += o.z;
        const float a2 = (fl & 1) ? 1.f : 0.f, asp = (fl & 2) ? 1.f : 0.f;
        a.wdir[d_out] = (act_tanh<true>(s0) + act_tanh<true>(s1) * a2 + act_tanh<true>(s2) * asp) / 3.0f;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(8 + wq) : "memory");       // spart may be overwritten by the next tile
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");      // warps 18, 19: their registers go to the build warps
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();           // the peer may still be read by / written from the leader's MMAs until here
  if (warp == 17) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm<512>(tmem_base);
  }
}

// Modulate vectors of the coordinate heads of ALL blocks as bf16, once per denoiser call: cmod[(l * B + b) * 512 + c] =
// shift[c] (c < 256) | 1 + scale[c - 256], from the fp32 adaLN table (ADA_COORD columns of block l).
__global__ void __launch_bounds__(256) k_coord_mod(int B, const float* __restrict__ ada, bf16* __restrict__ cmod) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, l = blockIdx.y, t = threadIdx.x;
  const float* ar = ada + static_cast<size_t>(b) * ADA_LD + l * ADA_BLK + ADA_COORD;
  bf16* o = cmod + (static_cast<size_t>(l) * B + b) * 512;
  o[t] = __float2bfloat16_rn(ar[t]);
  o[256 + t] = __float2bfloat16_rn(1.0f + ar[256 + t]);
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

int coord_head_launch(DsContext* ctx, const Plan& plan, const void* X, const void* ab, const void* cmod_l, const uint8_t* pflags,
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
  a.cmod = reinterpret_cast<const bf16*>(cmod_l);
  a.bc1 = bc1;
  a.wc2 = wc2;
  a.pflags = pflags;
  a.pair_info = plan.pair_info;
  a.pair_rows = plan.pair_rows;
  a.n_atoms = plan.n_atoms;
  a.poff = plan.poff;
  a.wdir = wdir;
  a.Mp = plan.Mp;
  a.dbg = nullptr;
  if (const char* e = getenv("DS_COORD_DBG")) a.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));   // developer aid: device pointer
  const int tiles = (plan.Mp + TM - 1) / TM;
  const int max_cl = ctx->num_sms / 2;
  const int ncl = tiles < max_cl ? tiles : max_cl;
  ds_launch(coord_head_kernel, dim3(2 * ncl), dim3(kThreads), kSmem, s, tmX, tmWe, tmWc1, tmAB, a);
  DS_CUDA_CHECK(cudaGetLastError());
  ctx->launch_count++;
  return DS_OK;
}

int coord_mod_launch(DsContext* ctx, int B, int n_blocks, const float* ada, void* cmod, cudaStream_t s) {
  ds_launch(k_coord_mod, dim3(B, n_blocks), dim3(256), 0, s, B, ada, reinterpret_cast<bf16*>(cmod));
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
