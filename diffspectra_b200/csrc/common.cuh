// Shared device helpers and the internal (C++) interface between the translation units of
// libdiffspectra_b200.so.  Nothing here is part of the C-ABI (see include/diffspectra_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

// ----------------------------------------------------------------------------- error codes (mirrored in the header)
#ifndef DS_OK
#define DS_OK 0
#define DS_ERR_INVALID (-1)
#define DS_ERR_CUDA (-2)
#define DS_ERR_MISSING_PARAM (-3)
#define DS_ERR_UNSUPPORTED (-4)
#define DS_ERR_WORKSPACE (-5)
#endif

#define DS_CUDA_CHECK(expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ds_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return DS_ERR_CUDA;                                                                         \
    }                                                                                             \
  } while (0)

#define DS_CHECK(cond, code, ...) \
  do {                            \
    if (!(cond)) {                \
      ds_set_error(__VA_ARGS__);  \
      return (code);              \
    }                             \
  } while (0)

#define DS_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != DS_OK) return _r; \
  } while (0)

void ds_set_error(const char* fmt, ...);

// ----------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the step graph CAN be launched with programmaticStreamSerialization: it may become resident while
// its predecessor drains, runs its prologue, and blocks in pdl_wait() until the predecessor has completed and its
// memory is visible.  pdl_trigger() at the top lets the successor do the same.  (A kernel launched without the
// attribute, or first in the stream, sees both as no-ops.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef DS_PDL_EARLY_TRIGGER
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_trigger() {}     // implicit trigger at grid completion: only the launch is pre-staged
#endif
// In-stream kernel timing (ds_profile_begin / ds_profile_end): every launch is bracketed by two events on its stream.
struct DsProfRec { const void* fn; long long tag; cudaEvent_t e0, e1; };
struct DsProf { bool on = false; long long tag = 0; std::vector<DsProfRec> recs; };   // tag: set by the GEMM launcher (shape)
extern DsProf g_ds_prof;
extern bool g_ds_use_pdl;      // DS_PDL=1 enables the launch attribute (off by default, see api_core.cu)
template <typename... KArgs, typename... Args>
inline cudaError_t ds_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_ds_use_pdl ? 1 : 0;
  if (!g_ds_prof.on) return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  DsProfRec r{reinterpret_cast<const void*>(kernel), g_ds_prof.tag, nullptr, nullptr};
  g_ds_prof.tag = 0;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, s);
  const cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  cudaEventRecord(r.e1, s);
  g_ds_prof.recs.push_back(r);
  return err;
}

// ----------------------------------------------------------------------------- model dimensions (QM9S config)
// configs/diffspectra_qm9s.py:53-76 — the kernels are specialised for these.
constexpr int D_NODE = 256;    // model.nf
constexpr int D_EDGE = 64;     // nf / 4
constexpr int D_TIME = 1024;   // nf * 4
constexpr int N_LAYERS = 8;
constexpr int N_HEADS = 16;
constexpr int N_XHEADS = 2;                 // adjacency heads
constexpr int N_SUB = N_HEADS - N_XHEADS;   // 14 learned heads
constexpr int C_SUB = 18;                   // 256 / 14
constexpr int C_HEAD = 16;
constexpr int QK_DIM = N_SUB * C_SUB;       // 252
// Column order of q, k (inside qkv) and e0 (inside e01): channel d of head h lives at head_perm(h * C_SUB + d), i.e. the 14
// heads' channel PAIRS are interleaved ([d/2][h][d%2]).  The logits are sums over channels of q*k*e0 per head, so any
// common permutation is exact; this one makes the 14 lanes that work on one (source, target) pair read 56 contiguous
// bytes per load instead of 14 words 36 bytes apart (2 sectors instead of 16 through L1 per request).  Applied once,
// to the rows of lin_query / lin_key / lin_edge0 (and the biases), when the weights are packed.
__host__ __device__ constexpr int head_perm(int c) { return (((c % C_SUB) >> 1) * N_SUB + c / C_SUB) * 2 + (c & 1); }
constexpr int QK_PAIR_STRIDE = 2 * N_SUB;   // 28 elements between consecutive channel pairs of one head
constexpr int QKV_LD = 768;                 // q[0,252) pad, k[256,508) pad, v[512,768)
constexpr int E01_LD = 512;                 // e0[0,252) pad, e1[256,512)
constexpr int MAX_ATOMS = 64;               // stress config; QM9S has <= 29

// adaLN table layout: one row per molecule (fp32)
constexpr int ADA_NODE = 0;      // nsh1 nsc1 ng1 nsh2 nsc2 ng2  (6 x 256)
constexpr int ADA_EDGE = 1536;   // esh1 esc1 eg1 esh2 esc2 eg2  (6 x 64)
constexpr int ADA_COORD = 1920;  // csh csc                      (2 x 256)
constexpr int ADA_RBF = 2432;    // scale, shift
constexpr int ADA_BLK = 2440;    // per-block stride (padded)
constexpr int ADA_ROOT_RBF = N_LAYERS * ADA_BLK;   // 19520
constexpr int ADA_LD = 19584;                      // padded row length

// ----------------------------------------------------------------------------- activation / dtype tags
// ACT_SILU_HALF: the input is h = x/2 (the producing Linear was packed with 0.5 W, 0.5 b, exact in bf16):
// SiLU(x) = x sigmoid(x) = h + h tanh(h) — one MUFU + one FFMA instead of two multiplies more.
// ACT_TANH_MIX (bf16 GEMM epilogues only): tanh with every second column pair on the FMA pipe (tanh_poly2) instead of MUFU.TANH.
enum DsAct { ACT_NONE = 0, ACT_SILU = 1, ACT_TANH = 2, ACT_GELU = 3, ACT_SILU_HALF = 4, ACT_TANH_MIX = 5 };
enum DsDType { DT_F32 = 0, DT_BF16 = 1 };

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ bf16 from_f32<bf16>(float v) {
  return __float2bfloat16_rn(v);
}

// kFast = bf16 production mode (approximate SFU maths), !kFast = fp32 validation mode (libm-accurate).
template <bool kFast>
__device__ __forceinline__ float act_tanh(float x) {
  if (kFast) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
  return tanhf(x);
}
template <bool kFast>
__device__ __forceinline__ float act_exp(float x) {
  return kFast ? __expf(x) : expf(x);
}
template <bool kFast>
__device__ __forceinline__ float act_silu(float x) {
  if (kFast) return x * (0.5f * act_tanh<true>(0.5f * x) + 0.5f);   // x*sigmoid(x), one MUFU
  return x / (1.0f + expf(-x));
}
template <bool kFast>
__device__ __forceinline__ float act_silu_half(float h) {   // SiLU(2h)
  if (kFast) return fmaf(h, act_tanh<true>(h), h);
  return (2.0f * h) / (1.0f + expf(-2.0f * h));
}
__device__ __forceinline__ float act_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <bool kFast>
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_SILU: return act_silu<kFast>(x);
    case ACT_TANH:
    case ACT_TANH_MIX: return act_tanh<kFast>(x);       // outside the tcgen05 STORE epilogue the mix is plain tanh
    case ACT_GELU: return act_gelu(x);
    case ACT_SILU_HALF: return act_silu_half<kFast>(x);
    default: return x;
  }
}

// ----------------------------------------------------------------------------- packed fp32 (sm_100: FFMA2 / FADD2 / FMUL2)
// Two IEEE fp32 operations per issued instruction (same rounding as the scalar forms, lane by lane).  The CUDA-core
// kernels here are bound by issue slots, not by the fp32 pipe, so halving the instruction count of their FMA runs pays.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

constexpr float TANH_C0 = 9.968600869e-01f, TANH_C1 = -3.149698377e-01f, TANH_C2 = 1.002266556e-01f, TANH_C3 = -2.373802848e-02f,
                TANH_C4 = 3.820235375e-03f, TANH_C5 = -3.979499161e-04f, TANH_C6 = 2.547285476e-05f, TANH_C7 = -9.067794053e-07f,
                TANH_C8 = 1.370746450e-08f;
// tanh on the FMA pipe: odd minimax polynomial x P(x^2) with 9 terms on [-3.75, 3.75] (input clamped), max abs error 6.0e-4 over the
// whole real line (MUFU.TANH: ~5e-4).  MUFU.TANH sustains ~8 results per clock per SM, a tanh epilogue over 83 M elements is
// bound by it (37 us); 2 packed multiplies + 8 packed FMAs per PAIR run on the otherwise idle FMA pipe, so an epilogue that
// sends every second pair here halves its MUFU time.  (The same split the FlashAttention-4 softmax uses for exp2.)
__device__ __forceinline__ float2 tanh_poly2(float2 x) {
  x.x = fminf(fmaxf(x.x, -3.75f), 3.75f);
  x.y = fminf(fmaxf(x.y, -3.75f), 3.75f);
  const float2 t = fmul2(x, x);
  float2 p = make_float2(TANH_C8, TANH_C8);
  p = ffma2(p, t, make_float2(TANH_C7, TANH_C7));
  p = ffma2(p, t, make_float2(TANH_C6, TANH_C6));
  p = ffma2(p, t, make_float2(TANH_C5, TANH_C5));
  p = ffma2(p, t, make_float2(TANH_C4, TANH_C4));
  p = ffma2(p, t, make_float2(TANH_C3, TANH_C3));
  p = ffma2(p, t, make_float2(TANH_C2, TANH_C2));
  p = ffma2(p, t, make_float2(TANH_C1, TANH_C1));
  p = ffma2(p, t, make_float2(TANH_C0, TANH_C0));
  return fmul2(p, x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------- molecule plan (packed ragged layout)
// node_info[m] = mol << 6 | i          pair_info[p] = mol << 12 | i << 6 | j   (i < j)
struct Plan {
  int B, N;              // molecules, padded atom count of the dense interface
  int Mn, Mp;            // total atoms, total unordered pairs
  const int* n_atoms;    // [B]
  const int* noff;       // [B+1]
  const int* poff;       // [B+1]
  const uint32_t* node_info;   // [Mn]
  const uint32_t* pair_info;   // [Mp]
  // [2*Mp] directed edges, source-major (d = 2*poff[mol] + r*(n-1) + c - (c > r)):
  // x = pair row, y = atom row of the source r, z = atom row of the target c, w = molecule
  const int4* dir_info;
  const uint32_t* dir_mol;     // [2*Mp] molecule of each directed edge (row_info form for the GEMM epilogues)
  const int2* pair_rows;       // [Mp] atom rows (i, j) of each unordered pair
  const int* mol_order;        // [B] molecules by descending atom count (launch order of the per-molecule kernels)
  const int* node_order;       // [Mn] atom rows in that molecule order (launch order of the warp-per-atom kernels)
  // the same two orders with everything a CTA / warp needs to start in ONE 16-byte load (instead of a chain of three):
  const int4* mol_launch;      // [B]  (molecule, n_atoms, noff, poff) of the k-th molecule in launch order
  const int4* atom_launch;     // [Mn] (atom row, molecule, n_atoms << 8 | index in molecule, poff) of the k-th atom in launch order
};
__device__ __forceinline__ int pair_index(int n, int i, int j) {   // i < j < n, row-major upper triangle
  return i * n - (i * (i + 1)) / 2 + (j - i - 1);
}

// ----------------------------------------------------------------------------- GEMM  out = act(A W^T + bias + addmat)
enum GemmMode { GEMM_STORE = 0, GEMM_LNMOD = 1, GEMM_RESGATE = 2, GEMM_COORD = 3, GEMM_EHEAD = 4 };

struct GemmDesc {
  const void* A = nullptr;      // [M, K] row-major, leading dim lda (elements); dtype a_dtype
  const void* W = nullptr;      // [N, K] row-major, leading dim ldw; same dtype as A
  const float* bias = nullptr;  // [N] or null
  const float* addmat = nullptr;   // [M, ldadd] fp32 added before the activation, or null
  void* out = nullptr;          // [M, N] row-major, leading dim ldo; dtype out_dtype
  int M = 0, N = 0, K = 0;
  int lda = 0, ldw = 0, ldo = 0, ldadd = 0;
  int a_dtype = DT_F32, out_dtype = DT_F32;
  int act = ACT_NONE;
  // fused epilogues of the tcgen05 kernel (gemm_tc.cu)
  int mode = GEMM_STORE;
  const uint32_t* row_info = nullptr;   // row -> molecule: row_info[row] >> info_shift
  int info_shift = 0;
  const float* ada = nullptr;           // adaLN table pre-offset to the block; row stride ADA_LD
  int off_a = 0, off_b = 0;             // LNMOD: shift / scale offsets; RESGATE: gate offset
  const float* resid = nullptr;         // RESGATE residual stream [M, ldres] fp32
  int ldres = 0;
  void* out2 = nullptr;                 // RESGATE bf16 copy [M, ldo2]
  int ldo2 = 0;
  int split_n = 0;                      // STORE (bf16, TMA out): output columns >= split_n go to out2 (ld ldo2) instead, 0 = off
  const float* wc2 = nullptr;           // COORD: coord_mlp.2 [3,256]
  const uint8_t* pflags = nullptr;      // COORD: adjacency bits per pair
  float* wdir = nullptr;                // COORD: output per directed edge
};

struct DsContext;
int gemm_simt_launch(const GemmDesc& g, bool fast_math, cudaStream_t s);
int gemm_tc_launch(DsContext* ctx, const GemmDesc& g, cudaStream_t s);   // bf16 in, tcgen05
int gemm_tc_init(DsContext* ctx);
// 2-D SWIZZLE_128B tensor map over a row-major [rows, cols] matrix with leading dimension ld (elements)
int ds_make_tmap_2d(DsContext* ctx, CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_cols,
                    int box_rows, bool f32);
