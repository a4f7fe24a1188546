// SpecFormer spectral encoder + cond_lin, run ONCE per molecule and cached across all sampling steps
// (the reference recomputes it every denoiser call, models/dmt.py:348-350).
// Reference: models/specformer.py:77-120 (patching), :167-200 (TSTiEncoder), :279-309 (post-norm layer with
// BatchNorm1d in eval mode), :345-425 (multi-head attention with residual pre-softmax scores), :457-470 (head).
#include "kernels.cuh"

namespace {

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
#define LAUNCH_CHECK(ctx)              \
  do {                                 \
    DS_CUDA_CHECK(cudaGetLastError()); \
    (ctx)->launch_count++;             \
  } while (0)

constexpr int SD = 128;    // d_model
constexpr int SH = 16;     // heads
constexpr int SK = 8;      // d_k = d_v

// z[b, q0+q, :] = W_P patch(b, q) + b_P + W_pos[q]      (specformer.py:96,174-189)
template <typename AT>
__global__ void __launch_bounds__(128) k_patch_embed(const float* __restrict__ spec, int L, int patch_len, int stride, int P,
                                                     int q0, int Q, const float* __restrict__ w, const float* __restrict__ b,
                                                     const float* __restrict__ wpos, float* __restrict__ z,
                                                     AT* __restrict__ zb) {
  __shared__ float patch[64];
  const int q = blockIdx.x, bi = blockIdx.y, t = threadIdx.x;
  if (t < patch_len) patch[t] = spec[static_cast<size_t>(bi) * L + q * stride + t];
  __syncthreads();
  float acc = b[t];
  for (int k = 0; k < patch_len; ++k) acc = fmaf(w[t * patch_len + k], patch[k], acc);
  acc += wpos[q * SD + t];
  const size_t o = (static_cast<size_t>(bi) * Q + q0 + q) * SD + t;
  z[o] = acc;
  zb[o] = from_f32<AT>(acc);
}

// One CTA per (molecule, head): K and V of the head live in shared memory (rows padded to 12 floats so that the
// 128-bit reads of consecutive keys are bank-conflict free); each warp takes queries i = w, w+8, ...:
// s_j = scale * q_i . k_j + prev[b,h,i,j]; scores <- s; a = softmax_j(s); out[b, i, h*8:(h+1)*8] = sum_j a_j v_j
// (specformer.py:395-419).  HBM traffic = the residual score tensor (read + write), everything else stays on chip.
constexpr int SPEC_MAXQ = 352;
template <typename AT, bool kFast>
__global__ void __launch_bounds__(256) k_spec_attn(const float* __restrict__ qkv, float* __restrict__ scores, int first,
                                                   const float* __restrict__ scale_p, int Q, AT* __restrict__ out) {
  __shared__ __align__(16) float sk[SPEC_MAXQ * 12];
  __shared__ __align__(16) float sv[SPEC_MAXQ * 12];
  const int h = blockIdx.x % SH, bi = blockIdx.x / SH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = scale_p[0];
  for (int idx = threadIdx.x; idx < Q * 2; idx += 256) {
    const int j = idx >> 1, half = idx & 1;
    const float* src = qkv + (static_cast<size_t>(bi) * Q + j) * 384 + h * SK + half * 4;
    *reinterpret_cast<float4*>(sk + j * 12 + half * 4) = *reinterpret_cast<const float4*>(src + 128);
    *reinterpret_cast<float4*>(sv + j * 12 + half * 4) = *reinterpret_cast<const float4*>(src + 256);
  }
  __syncthreads();
  constexpr int MAXJ = SPEC_MAXQ / 32;    // 11
  for (int i = warp; i < Q; i += 8) {
    const float* qr = qkv + (static_cast<size_t>(bi) * Q + i) * 384 + h * SK;
    const float4 q0 = *reinterpret_cast<const float4*>(qr), q1 = *reinterpret_cast<const float4*>(qr + 4);
    float* sr = scores + (static_cast<size_t>(bi) * SH + h) * Q * Q + static_cast<size_t>(i) * Q;
    float svv[MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int it = 0; it < MAXJ; ++it) {
      const int j = lane + 32 * it;
      svv[it] = -INFINITY;
      if (j < Q) {
        const float4 k0 = *reinterpret_cast<const float4*>(sk + j * 12), k1 = *reinterpret_cast<const float4*>(sk + j * 12 + 4);
        float dot = q0.x * k0.x;
        dot = fmaf(q0.y, k0.y, dot); dot = fmaf(q0.z, k0.z, dot); dot = fmaf(q0.w, k0.w, dot);
        dot = fmaf(q1.x, k1.x, dot); dot = fmaf(q1.y, k1.y, dot); dot = fmaf(q1.z, k1.z, dot); dot = fmaf(q1.w, k1.w, dot);
        float sc = dot * scale;
        if (!first) sc += sr[j];
        sr[j] = sc;
        svv[it] = sc;
        mx = fmaxf(mx, sc);
      }
    }
    mx = warp_max(mx);
    float den = 0.f;
    float o[SK];
#pragma unroll
    for (int d = 0; d < SK; ++d) o[d] = 0.f;
#pragma unroll
    for (int it = 0; it < MAXJ; ++it) {
      const int j = lane + 32 * it;
      if (j < Q) {
        const float e = act_exp<kFast>(svv[it] - mx);
        den += e;
        const float4 v0 = *reinterpret_cast<const float4*>(sv + j * 12), v1 = *reinterpret_cast<const float4*>(sv + j * 12 + 4);
        o[0] = fmaf(e, v0.x, o[0]); o[1] = fmaf(e, v0.y, o[1]); o[2] = fmaf(e, v0.z, o[2]); o[3] = fmaf(e, v0.w, o[3]);
        o[4] = fmaf(e, v1.x, o[4]); o[5] = fmaf(e, v1.y, o[5]); o[6] = fmaf(e, v1.z, o[6]); o[7] = fmaf(e, v1.w, o[7]);
      }
    }
    den = warp_sum(den);
#pragma unroll
    for (int d = 0; d < SK; ++d) o[d] = warp_sum(o[d]);
    if (lane < SK) {
      float v = o[0];
#pragma unroll
      for (int d = 1; d < SK; ++d) v = (lane == d) ? o[d] : v;
      out[(static_cast<size_t>(bi) * Q + i) * SD + h * SK + lane] = from_f32<AT>(v / den);
    }
  }
}

// z = BatchNorm_eval(z + delta)  over the channel dim (specformer.py:292-294,303-306)
template <typename AT>
__global__ void k_add_bn(float* __restrict__ z, const float* __restrict__ delta, const float* __restrict__ bn, size_t total,
                         AT* __restrict__ zb) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx & (SD - 1);
  const float v = z[idx] + delta[idx];
  const float y = (v - bn[256 + c]) / sqrtf(bn[384 + c] + 1e-5f) * bn[c] + bn[128 + c];
  z[idx] = y;
  zb[idx] = from_f32<AT>(y);
}

// out_norm: LayerNorm(256) with affine, eps 1e-5 (specformer.py:67,119); one warp per row
template <typename AT>
__global__ void __launch_bounds__(256) k_ln_affine256(const float* __restrict__ x, const float* __restrict__ wb, int rows,
                                                      AT* __restrict__ out) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = x[static_cast<size_t>(r) * 256 + lane + 32 * i];
    s += v[i];
  }
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] -= mean;
    q += v[i] * v[i];
  }
  const float is = 1.0f / sqrtf(warp_sum(q) * (1.0f / 256.0f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    out[static_cast<size_t>(r) * 256 + c] = from_f32<AT>(v[i] * is * wb[c] + wb[256 + c]);
  }
}

template <typename AT, bool kFast>
int spec_impl(DsContext* ctx, const PackedWeights& pw, const float* const* spectra, int B, float* ctx_out, void* workspace,
              size_t ws_bytes, cudaStream_t s) {
  static const int kLen[3] = {701, 3501, 3501}, kPatch[3] = {20, 50, 50}, kStride[3] = {10, 25, 25};
  const int AD = kFast ? DT_BF16 : DT_F32;
  const int Q = pw.q_len;
  // chunk the batch so the residual-score tensor [Bc,16,Q,Q] stays bounded (<= ~1 GB)
  int Bc = static_cast<int>((size_t(1) << 30) / (static_cast<size_t>(SH) * Q * Q * 4));
  if (Bc < 1) Bc = 1;
  if (Bc > B) Bc = B;
  Arena a{static_cast<uint8_t*>(workspace), 0, ws_bytes, false};
  SpecWs w;
  spec_ws_carve(a, w, Bc, Q, kFast);
  DS_CHECK(a.off <= ws_bytes, DS_ERR_WORKSPACE, "specformer: workspace too small (%zu < %zu)", ws_bytes, a.off);
  for (int b0 = 0; b0 < B; b0 += Bc) {
    const int nb = (B - b0 < Bc) ? (B - b0) : Bc;
    const int rows = nb * Q;
    int q0 = 0;
    for (int si = 0; si < pw.n_spec; ++si) {
      const int ty = pw.spec_type[si];
      dim3 grid(pw.patch_num[si], nb);
      k_patch_embed<AT><<<grid, 128, 0, s>>>(spectra[si] + static_cast<size_t>(b0) * kLen[ty], kLen[ty], kPatch[ty],
                                             kStride[ty], pw.patch_num[si], q0, Q, pw.wp_w[si], pw.wp_b[si], pw.w_pos[si],
                                             w.z, reinterpret_cast<AT*>(w.zb));
      LAUNCH_CHECK(ctx);
      q0 += pw.patch_num[si];
    }
    for (int l = 0; l < 3; ++l) {
      const SpecLayerWeights& sl = pw.sl[l];
      DS_TRY(linear(ctx, w.zb, SD, sl.wqkv, SD, sl.bqkv, nullptr, 0, w.qkv, 384, DT_F32, rows, 384, SD, ACT_NONE, s));
      k_spec_attn<AT, kFast><<<nb * SH, 256, 0, s>>>(w.qkv, w.scores, l == 0, sl.scale, Q, reinterpret_cast<AT*>(w.att));
      LAUNCH_CHECK(ctx);
      DS_TRY(linear(ctx, w.att, SD, sl.wo, SD, sl.bo, nullptr, 0, w.o, SD, DT_F32, rows, SD, SD, ACT_NONE, s));
      const size_t total = static_cast<size_t>(rows) * SD;
      k_add_bn<AT><<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w.z, w.o, sl.bn1, total, reinterpret_cast<AT*>(w.zb));
      LAUNCH_CHECK(ctx);
      DS_TRY(linear(ctx, w.zb, SD, sl.wf0, SD, sl.bf0, nullptr, 0, w.f, 256, AD, rows, 256, SD, ACT_GELU, s));
      DS_TRY(linear(ctx, w.f, 256, sl.wf3, 256, sl.bf3, nullptr, 0, w.o, SD, DT_F32, rows, SD, 256, ACT_NONE, s));
      k_add_bn<AT><<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w.z, w.o, sl.bn2, total, reinterpret_cast<AT*>(w.zb));
      LAUNCH_CHECK(ctx);
    }
    // Flatten_Head (specformer.py:457-470): [nb, Q*128] x [256, Q*128]^T, then out_norm, then cond_lin (dmt.py:350)
    DS_TRY(linear(ctx, w.zb, Q * SD, pw.head_w, Q * SD, pw.head_b, nullptr, 0, w.head, 256, DT_F32, nb, 256, Q * SD, ACT_NONE, s));
    k_ln_affine256<AT><<<cdiv(nb, 8), 256, 0, s>>>(w.head, pw.out_norm, nb, reinterpret_cast<AT*>(w.hb));
    LAUNCH_CHECK(ctx);
    DS_TRY(linear(ctx, w.hb, 256, pw.cond_w, 256, pw.cond_b, nullptr, 0, ctx_out + static_cast<size_t>(b0) * D_TIME, D_TIME,
                  DT_F32, nb, D_TIME, 256, ACT_NONE, s));
  }
  return DS_OK;
}

}  // namespace

size_t spec_ws_carve(Arena& a, SpecWs& w, int Bc, int Q, bool bf) {
  const size_t es = bf ? 2 : 4;
  const size_t rows = static_cast<size_t>(Bc) * Q;
  const size_t start = a.off;
  w.z = static_cast<float*>(a.take(rows * SD * 4));
  w.zb = a.take(rows * SD * es);
  w.qkv = static_cast<float*>(a.take(rows * 384 * 4));
  w.scores = static_cast<float*>(a.take(static_cast<size_t>(Bc) * SH * Q * Q * 4));
  w.att = a.take(rows * SD * es);
  w.o = static_cast<float*>(a.take(rows * SD * 4));
  w.f = a.take(rows * 256 * es);
  w.head = static_cast<float*>(a.take(static_cast<size_t>(Bc) * 256 * 4));
  w.hb = a.take(static_cast<size_t>(Bc) * 256 * es);
  return a.off - start;
}

int specformer_ctx(DsContext* ctx, const PackedWeights& pw, const float* const* spectra, int B, float* ctx_out,
                   void* workspace, size_t ws_bytes, cudaStream_t s) {
  DS_CHECK(pw.valid, DS_ERR_INVALID, "specformer: weights not packed");
  DS_CHECK(B > 0, DS_ERR_INVALID, "specformer: empty batch");
  for (int i = 0; i < pw.n_spec; ++i) DS_CHECK(spectra[i] != nullptr, DS_ERR_INVALID, "specformer: spectrum %d is null", i);
  if (ds_is_bf16(ctx)) return spec_impl<bf16, true>(ctx, pw, spectra, B, ctx_out, workspace, ws_bytes, s);
  return spec_impl<float, false>(ctx, pw, spectra, B, ctx_out, workspace, ws_bytes, s);
}
