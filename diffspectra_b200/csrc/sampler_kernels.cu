// Sampler-side kernels: dense<->packed state conversion, the fused ancestral update (schedule coefficients,
// masked centre-of-mass-free Gaussian noise for coordinates, atom / bond channel update), initial noise,
// device-side Philox noise, and post_process.
//
// Reference: sampling.py:565-631 (AncestralSampler.sampling), :53-97 (post_process), models/utils.py:38-45,
// 67-106 (noise samplers), utils.py:71-105 (inverse scaler, factors 1,4,4,1, centered).
#include "kernels.cuh"

namespace {

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
#define LAUNCH_CHECK(ctx)              \
  do {                                 \
    DS_CUDA_CHECK(cudaGetLastError()); \
    (ctx)->launch_count++;             \
  } while (0)

// ----------------------------------------------------------------------------- Philox4x32-10 + Box-Muller
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  const float u2 = (static_cast<float>(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincosf(6.28318530717958647692f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}
constexpr uint32_t TAG_NODE = 0x4e4f4445u, TAG_EDGE = 0x45444745u;
// draw index `step` (step = -1 -> 0xffffffff: initial noise); atoms: call q of atom a = counter word a + 64 q (philox_node_quad)
__device__ __forceinline__ float2 philox_pair(unsigned long long seed, long long gid, int step, int i, int j) {
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(i * 64 + j), static_cast<uint32_t>(step),
                                           static_cast<uint32_t>(gid), TAG_EDGE ^ static_cast<uint32_t>(gid >> 32)), key);
  return box_muller(r.x, r.y);
}

// ----------------------------------------------------------------------------- dense <-> packed
__global__ void k_pack_nodes(Plan plan, const float* __restrict__ x, float* __restrict__ xs) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plan.Mn * 9) return;
  const int m = idx / 9, c = idx % 9;
  const uint32_t info = plan.node_info[m];
  const int mol = info >> 6, a = info & 63;
  xs[idx] = x[(static_cast<size_t>(mol) * plan.N + a) * 9 + c];
}
__global__ void k_pack_pairs(Plan plan, const float* __restrict__ ex, float* __restrict__ es) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plan.Mp * 2) return;
  const int p = idx >> 1, c = idx & 1;
  const uint32_t info = plan.pair_info[p];
  const int mol = info >> 12, i = (info >> 6) & 63, j = info & 63;
  es[idx] = ex[((static_cast<size_t>(mol) * plan.N + i) * plan.N + j) * 2 + c];
}
__global__ void k_unpack_nodes(Plan plan, const float* __restrict__ xs, float* __restrict__ x) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plan.B * plan.N * 9) return;
  const int c = idx % 9, a = (idx / 9) % plan.N, mol = idx / (9 * plan.N);
  x[idx] = (a < plan.n_atoms[mol]) ? xs[static_cast<size_t>(plan.noff[mol] + a) * 9 + c] : 0.f;
}
__global__ void k_unpack_pairs(Plan plan, const float* __restrict__ es, float* __restrict__ ex) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(plan.B) * plan.N * plan.N * 2;
  if (idx >= total) return;
  const int c = idx & 1;
  const int j = (idx >> 1) % plan.N, i = ((idx >> 1) / plan.N) % plan.N, mol = (idx >> 1) / (static_cast<size_t>(plan.N) * plan.N);
  const int n = plan.n_atoms[mol];
  float v = 0.f;
  if (i < n && j < n && i != j) {
    const int p = plan.poff[mol] + (i < j ? pair_index(n, i, j) : pair_index(n, j, i));
    v = es[static_cast<size_t>(p) * 2 + c];
  }
  ex[idx] = v;
}

// ----------------------------------------------------------------------------- ancestral update
// nodes: x <- c_x x + c_pred pred + sigma T eps,  eps = [CoM-free masked N(0,1)^3 | masked N(0,1)^6]      sampling.py:605-612
// One CTA of three warps per molecule: warp q owns channels [4q, 4q+4) (= the q-th Philox call of an atom: pos xyz + h0 |
// h1..h4 | h5), lanes are atoms (lane, lane + 32).  The centre of mass only involves warp 0.  (One warp per molecule did
// all nine channels of an atom in one thread: 1024 warps of ~600 dependent instructions each on 148 SMs, 26 us per step.)
__device__ __forceinline__ void philox_node_quad(unsigned long long seed, long long gid, int step, int a, int q, float (&z)[4]) {
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(a + 64 * q), static_cast<uint32_t>(step), static_cast<uint32_t>(gid),
                                           TAG_NODE ^ static_cast<uint32_t>(gid >> 32)), key);
  const float2 n0 = box_muller(r.x, r.y), n1 = box_muller(r.z, r.w);
  z[0] = n0.x; z[1] = n0.y; z[2] = n1.x; z[3] = n1.y;
}
__global__ void __launch_bounds__(96) k_sampler_nodes(Plan plan, float* __restrict__ xs, const float* __restrict__ pred,
                                                      float* __restrict__ xmean, StepRef sr, int step_host, NoiseSrc ns,
                                                      float temperature, int init_only) {
  pdl_trigger();
  pdl_wait();
  const int mol = blockIdx.x, q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = plan.n_atoms[mol], base = plan.noff[mol];
  const int step = sr.step ? *sr.step : step_host;
  const int nch = q == 2 ? 1 : 4;                 // channels of this warp: c = 4 q + k
  float z[2][4];
  float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
  for (int hq = 0; hq < 2; ++hq) {
    const int a = lane + 32 * hq;
#pragma unroll
    for (int k = 0; k < 4; ++k) z[hq][k] = 0.f;
    if (a < n) {
      if (ns.raw_pos) {
        const size_t so = static_cast<size_t>(init_only ? 0 : step - ns.raw_step_base) * plan.B * plan.N;
        const float* rp = ns.raw_pos + (so + static_cast<size_t>(mol) * plan.N + a) * 3;
        const float* rh = ns.raw_h + (so + static_cast<size_t>(mol) * plan.N + a) * 6;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = 4 * q + k;
          if (c < 9) z[hq][k] = c < 3 ? rp[c] : rh[c - 3];
        }
      } else {
        philox_node_quad(ns.seed, ns.gid_base + mol, init_only ? -1 : step + ns.philox_step_offset, a, q, z[hq]);
      }
      if (q == 0) { sx += z[hq][0]; sy += z[hq][1]; sz += z[hq][2]; }
    }
  }
  float mx = 0.f, my = 0.f, mz = 0.f;
  if (q == 0) {                                   // warp-uniform
    const float fn = static_cast<float>(n);
    mx = warp_sum(sx) / fn; my = warp_sum(sy) / fn; mz = warp_sum(sz) / fn;
  }
  float cx = 0.f, cp = 0.f, sg = 1.f;
  if (!init_only) {
    cx = sr.coef[step * 4 + 0]; cp = sr.coef[step * 4 + 1]; sg = sr.coef[step * 4 + 2];
  }
#pragma unroll
  for (int hq = 0; hq < 2; ++hq) {
    const int a = lane + 32 * hq;
    if (a >= n) continue;
    if (q == 0) { z[hq][0] -= mx; z[hq][1] -= my; z[hq][2] -= mz; }
    float* xr = xs + static_cast<size_t>(base + a) * 9 + 4 * q;
    if (init_only) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < nch) xr[k] = z[hq][k];
    } else {
      const float* pr = pred + static_cast<size_t>(base + a) * 9 + 4 * q;
      float* xm = xmean + static_cast<size_t>(base + a) * 9 + 4 * q;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nch) {
          const float mean = cx * xr[k] + cp * pr[k];
          xm[k] = mean;
          xr[k] = mean + (sg * z[hq][k]) * temperature;
        }
      }
    }
  }
}

// pairs: symmetric noise = strict lower triangle of randn[B,2,N,N] mirrored (models/utils.py:100-106)
__global__ void __launch_bounds__(256) k_sampler_pairs(Plan plan, float* __restrict__ es, const float* __restrict__ pred_e,
                                                       float* __restrict__ emean, StepRef sr, int step_host, NoiseSrc ns,
                                                       float temperature, int init_only) {
  pdl_trigger();
  pdl_wait();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= plan.Mp) return;
  const uint32_t info = plan.pair_info[p];
  const int mol = info >> 12, i = (info >> 6) & 63, j = info & 63;    // i < j
  const int step = sr.step ? *sr.step : step_host;
  float z0, z1;
  if (ns.raw_e) {
    const size_t NN = static_cast<size_t>(plan.N) * plan.N;
    const size_t so = static_cast<size_t>(init_only ? 0 : step - ns.raw_step_base) * plan.B * 2 * NN;
    const float* r = ns.raw_e + so + static_cast<size_t>(mol) * 2 * NN + static_cast<size_t>(j) * plan.N + i;   // row j > col i
    z0 = r[0];
    z1 = r[NN];
  } else {
    const float2 zz = philox_pair(ns.seed, ns.gid_base + mol, init_only ? -1 : step + ns.philox_step_offset, i, j);
    z0 = zz.x;
    z1 = zz.y;
  }
  float* er = es + static_cast<size_t>(p) * 2;
  if (init_only) {
    er[0] = z0;
    er[1] = z1;
    return;
  }
  const float cx = sr.coef[step * 4 + 0], cp = sr.coef[step * 4 + 1], sg = sr.coef[step * 4 + 2];
  const float m0 = cx * er[0] + cp * pred_e[p * 2 + 0];
  const float m1 = cx * er[1] + cp * pred_e[p * 2 + 1];
  emean[p * 2 + 0] = m0;
  emean[p * 2 + 1] = m1;
  er[0] = m0 + (sg * z0) * temperature;
  er[1] = m1 + (sg * z1) * temperature;
}

__global__ void k_step_inc(int* step) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) *step += 1;
}

// ----------------------------------------------------------------------------- post_process (sampling.py:53-97)
__global__ void k_post_nodes(Plan plan, const float* __restrict__ xs, float* __restrict__ pos, int* __restrict__ atom_type,
                             int* __restrict__ fc) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plan.B * plan.N) return;
  const int a = idx % plan.N, mol = idx / plan.N;
  float px = 0.f, py = 0.f, pz = 0.f;
  int ty = 0, q = 0;
  if (a < plan.n_atoms[mol]) {
    const float* xr = xs + static_cast<size_t>(plan.noff[mol] + a) * 9;
    px = xr[0]; py = xr[1]; pz = xr[2];                  // pos * 1 * mask
    float best = (xr[3] * 4.0f + 1.0f) / 2.0f;           // (atom_type * 4 + 1) / 2
    for (int c = 1; c < 5; ++c) {
      const float v = (xr[3 + c] * 4.0f + 1.0f) / 2.0f;
      if (v > best) { best = v; ty = c; }
    }
    q = static_cast<int>(rintf(xr[8] * 4.0f));           // round(fc * 4)  (half-to-even like torch.round)
  }
  pos[idx * 3 + 0] = px; pos[idx * 3 + 1] = py; pos[idx * 3 + 2] = pz;
  atom_type[idx] = ty;
  fc[idx] = q;
}
__global__ void k_post_pairs(Plan plan, const float* __restrict__ es, float* __restrict__ bond) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(plan.B) * plan.N * plan.N;
  if (idx >= total) return;
  const int j = idx % plan.N, i = (idx / plan.N) % plan.N, mol = idx / (static_cast<size_t>(plan.N) * plan.N);
  const int n = plan.n_atoms[mol];
  float v = 0.f;
  if (i < n && j < n && i != j) {
    const int p = plan.poff[mol] + (i < j ? pair_index(n, i, j) : pair_index(n, j, i));
    const float ex = (es[static_cast<size_t>(p) * 2 + 0] + 1.0f) / 2.0f;
    const float t = (es[static_cast<size_t>(p) * 2 + 1] + 1.0f) / 2.0f * 3.0f;
    const float exist = ex >= 0.5f ? 1.f : 0.f;
    const float order = t >= 2.5f ? 3.f : (t >= 1.5f ? 2.f : (t >= 0.5f ? 1.f : 0.f));
    v = exist * order;
  }
  bond[idx] = v;
}

// Molecule records straight from the reference-shaped means (sampling.py:12-32 mol_process + :53-97 post_process in one
// pass): record b = pos f32[R*3] | atom u8[R] | fc i8[R] | bond u8[R*R] | n u8, rec_bytes = 14 R + R^2 + 1 (1248 B at
// R = 29), R = rec_n >= N (rounds whose padded size N is smaller than the data set's max_node still emit full-size
// records).  This is the unit of the ONE all-gather of a sampling round and of the single D2H copy of the eval driver.
// One CTA per molecule; bytes are written individually (the record stride is not a multiple of 4 for every R).
__global__ void __launch_bounds__(256) k_molecule_records(Plan plan, const float* __restrict__ x, const float* __restrict__ ex,
                                                          uint8_t* __restrict__ rec, int R, int rec_bytes) {
  pdl_trigger();
  pdl_wait();
  const int mol = blockIdx.x, N = plan.N, n = plan.n_atoms[mol], t = threadIdx.x;
  uint8_t* r = rec + static_cast<size_t>(mol) * rec_bytes;
  const float* xm = x + static_cast<size_t>(mol) * N * 9;
  for (int idx = t; idx < R * 3; idx += 256) {
    const int a = idx / 3, c = idx - a * 3;
    const uint32_t u = __float_as_uint(a < n ? xm[a * 9 + c] : 0.f);
    r[idx * 4 + 0] = u & 0xff; r[idx * 4 + 1] = (u >> 8) & 0xff; r[idx * 4 + 2] = (u >> 16) & 0xff; r[idx * 4 + 3] = u >> 24;
  }
  uint8_t* ra = r + R * 12;
  for (int a = t; a < R; a += 256) {
    int ty = 0, q = 0;
    if (a < n) {
      const float* xr = xm + a * 9;
      float best = (xr[3] * 4.0f + 1.0f) / 2.0f;           // same arithmetic as k_post_nodes
      for (int c = 1; c < 5; ++c) {
        const float v = (xr[3 + c] * 4.0f + 1.0f) / 2.0f;
        if (v > best) { best = v; ty = c; }
      }
      q = static_cast<int>(rintf(xr[8] * 4.0f));
    }
    ra[a] = static_cast<uint8_t>(ty);
    ra[R + a] = static_cast<uint8_t>(static_cast<int8_t>(q));
  }
  uint8_t* rb = r + R * 14;
  const float* em = ex + static_cast<size_t>(mol) * N * N * 2;
  for (int idx = t; idx < R * R; idx += 256) {
    const int i = idx / R, j = idx - i * R;
    uint8_t v = 0;
    if (i < n && j < n && i != j) {
      const int lo = i < j ? i : j, hi = i < j ? j : i;     // the packed state keeps (lo, hi): same element as k_post_pairs
      const float* e = em + (static_cast<size_t>(lo) * N + hi) * 2;
      const float exv = (e[0] + 1.0f) / 2.0f;
      const float tv = (e[1] + 1.0f) / 2.0f * 3.0f;
      const int order = tv >= 2.5f ? 3 : (tv >= 1.5f ? 2 : (tv >= 0.5f ? 1 : 0));
      v = static_cast<uint8_t>(exv >= 0.5f ? order : 0);
    }
    rb[idx] = v;
  }
  if (t == 0) r[R * 14 + R * R] = static_cast<uint8_t>(n);
}

}  // namespace

int launch_molecule_records(DsContext* ctx, const Plan& plan, const float* x_mean, const float* edge_mean, uint8_t* rec,
                            int rec_n, int rec_bytes, cudaStream_t s) {
  ds_launch(k_molecule_records, dim3(plan.B), dim3(256), 0, s, plan, x_mean, edge_mean, rec, rec_n, rec_bytes);
  LAUNCH_CHECK(ctx);
  return DS_OK;
}

int launch_pack_dense(DsContext* ctx, const Plan& plan, const float* x, const float* ex, float* xs, float* es, cudaStream_t s) {
  if (x) {
    ds_launch(k_pack_nodes, dim3(cdiv(plan.Mn * 9, 256)), dim3(256), 0, s, plan, x, xs);
    LAUNCH_CHECK(ctx);
  }
  if (ex && plan.Mp > 0) {
    ds_launch(k_pack_pairs, dim3(cdiv(plan.Mp * 2, 256)), dim3(256), 0, s, plan, ex, es);
    LAUNCH_CHECK(ctx);
  }
  return DS_OK;
}

int launch_unpack_dense(DsContext* ctx, const Plan& plan, const float* xs, const float* es, float* x, float* ex, cudaStream_t s) {
  if (x) {
    ds_launch(k_unpack_nodes, dim3(cdiv(plan.B * plan.N * 9, 256)), dim3(256), 0, s, plan, xs, x);
    LAUNCH_CHECK(ctx);
  }
  if (ex) {
    const size_t total = static_cast<size_t>(plan.B) * plan.N * plan.N * 2;
    ds_launch(k_unpack_pairs, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, plan, es, ex);
    LAUNCH_CHECK(ctx);
  }
  return DS_OK;
}

int launch_sampler_step(DsContext* ctx, const Plan& plan, float* xs, float* es, const float* pred_x, const float* pred_e,
                        float* xmean, float* emean, StepRef sr, int step_host, const NoiseSrc& ns, float temperature,
                        cudaStream_t s) {
  ds_launch(k_sampler_nodes, dim3(plan.B), dim3(96), 0, s, plan, xs, pred_x, xmean, sr, step_host, ns, temperature, 0);
  LAUNCH_CHECK(ctx);
  if (plan.Mp > 0) {
    ds_launch(k_sampler_pairs, dim3(cdiv(plan.Mp, 256)), dim3(256), 0, s, plan, es, pred_e, emean, sr, step_host, ns, temperature, 0);
    LAUNCH_CHECK(ctx);
  }
  return DS_OK;
}

int launch_init_noise(DsContext* ctx, const Plan& plan, float* xs, float* es, const NoiseSrc& ns, cudaStream_t s) {
  StepRef sr{nullptr, nullptr};
  ds_launch(k_sampler_nodes, dim3(plan.B), dim3(96), 0, s, plan, xs, nullptr, nullptr, sr, 0, ns, 1.0f, 1);
  LAUNCH_CHECK(ctx);
  if (plan.Mp > 0) {
    ds_launch(k_sampler_pairs, dim3(cdiv(plan.Mp, 256)), dim3(256), 0, s, plan, es, nullptr, nullptr, sr, 0, ns, 1.0f, 1);
    LAUNCH_CHECK(ctx);
  }
  return DS_OK;
}

int launch_step_inc(DsContext* ctx, int* step, cudaStream_t s) {
  ds_launch(k_step_inc, dim3(1), dim3(32), 0, s, step);
  LAUNCH_CHECK(ctx);
  return DS_OK;
}

int launch_post_process(DsContext* ctx, const Plan& plan, const float* xs, const float* es, float* pos, int* atom_type,
                        int* fc, float* bond, cudaStream_t s) {
  ds_launch(k_post_nodes, dim3(cdiv(plan.B * plan.N, 256)), dim3(256), 0, s, plan, xs, pos, atom_type, fc);
  LAUNCH_CHECK(ctx);
  const size_t total = static_cast<size_t>(plan.B) * plan.N * plan.N;
  ds_launch(k_post_pairs, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, plan, es, bond);
  LAUNCH_CHECK(ctx);
  return DS_OK;
}
