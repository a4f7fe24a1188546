// DMT_WO_EQ — the non-equivariant ablation (models/dmt_wo_eq.py:646-937) on the packed ragged layout.
// Textually included inside the anonymous namespace of dmt_kernels.cu (it reuses that file's helpers and the root /
// head kernels of DMT).  Differences from DMT that shape the kernels (last paragraph):
//   * edges are NOT symmetric after the first block (node2edge_lin acts on cat[h_r, h_c], :603-609), so every edge
//     tensor lives per DIRECTED edge, source-major (row d = 2*poff[mol] + r*(n-1) + c - (c > r) = PyG's edge order);
//   * attention (TransLayerOptimV2, :207-259): q of the target, (k, v) of the source PLUS per-edge (ek, ev), 16 heads
//     of 16 channels, no adjacency heads, an output projection;
//   * FFNs use GELU(erf); the FFN residual base is the un-normalised sum (:587-600, :615-626);
//   * no coordinate update: positions come from pos_pred_mlp on the concatenated atom features (:709-717, :915).
// node2edge_lin is hoisted per atom (two 256 -> 64 projections added per directed edge), the adaLN projections per
// molecule, exactly as in DMT.

// NodeEmbed hidden layer: GELU(x_linear([h | cond_h]) + pos_linear(pos))   (:638-643)
template <typename AT>
__global__ void __launch_bounds__(256) k_wo_root_nodes(int Mn, const float* __restrict__ xs, const float* __restrict__ cond,
                                                       const float* __restrict__ wx, const float* __restrict__ bx,
                                                       const float* __restrict__ wp, const float* __restrict__ bp,
                                                       AT* __restrict__ hid) {
  pdl_trigger();
  pdl_wait();
  __shared__ float in[15];
  const int m = blockIdx.x, t = threadIdx.x;
  if (t < 6) in[t] = xs[m * 9 + 3 + t];
  else if (t < 12) in[t] = cond ? cond[m * 9 + 3 + (t - 6)] : 0.f;
  else if (t < 15) in[t] = xs[m * 9 + (t - 12)];
  __syncthreads();
  for (int o = t; o < 512; o += 256) {
    float a = bx[o], b = bp[o];
#pragma unroll
    for (int k = 0; k < 12; ++k) a = fmaf(wx[o * 12 + k], in[k], a);
#pragma unroll
    for (int k = 0; k < 3; ++k) b = fmaf(wp[o * 3 + k], in[12 + k], b);
    hid[static_cast<size_t>(m) * 512 + o] = from_f32<AT>(act_gelu(a + b));
  }
}

// activation-dtype copies of a fp32 [rows, W] stream: dst0 (ld0) and optionally dst1 (ld1); 4 channels per thread
template <typename AT>
__global__ void k_wo_copy(int rows, int W, const float* __restrict__ src, AT* __restrict__ dst0, int ld0, AT* __restrict__ dst1,
                          int ld1) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = W / 4;
  if (idx >= rows * per) return;
  const int r = idx / per, c = (idx % per) * 4;
  const float4 v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * W + c);
  store4<AT>(dst0 + static_cast<size_t>(r) * ld0 + c, v.x, v.y, v.z, v.w);
  if (dst1) store4<AT>(dst1 + static_cast<size_t>(r) * ld1 + c, v.x, v.y, v.z, v.w);
}

// the root edge embedding is symmetric: expand per-pair rows to the two directed edges
template <typename AT>
__global__ void k_wo_expand_root(int Md, const int4* __restrict__ dir_info, const float* __restrict__ e_pair,
                                 float* __restrict__ e, AT* __restrict__ ehid) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Md * 16) return;
  const int d = idx >> 4, c = (idx & 15) * 4;
  const int p = dir_info[d].x;
  const float4 v = *reinterpret_cast<const float4*>(e_pair + static_cast<size_t>(p) * 64 + c);
  *reinterpret_cast<float4*>(e + static_cast<size_t>(d) * 64 + c) = v;
  store4<AT>(ehid + static_cast<size_t>(d) * 192 + c, v.x, v.y, v.z, v.w);
}

// ea = modulate(norm1_edge(e), esh1, esc1) per directed edge   (:519-521)
template <typename AT, bool kFast>
__global__ void __launch_bounds__(256) k_wo_dir_ln1(int Md, const uint32_t* __restrict__ dir_mol, const float* __restrict__ e,
                                                    const float* __restrict__ ada, int l, AT* __restrict__ ea) {
  pdl_trigger();
  pdl_wait();
  const int d = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (d >= Md) return;
  const float* ar = ada + static_cast<size_t>(dir_mol[d]) * ADA_LD + l * ADA_BLK + ADA_EDGE;
  const float2 v = *reinterpret_cast<const float2*>(e + static_cast<size_t>(d) * 64 + 2 * lane);
  float v0 = v.x, v1 = v.y;
  ln64_mod<kFast>(v0, v1, ar + 0, ar + 64, lane);
  store2<AT>(ea + static_cast<size_t>(d) * 64 + 2 * lane, v0, v1);
}

// TransLayerOptimV2 (:207-259): one CTA per (molecule, group of ATT_G targets).
//   logit[r -> c, h] = q[c,h,:] . (k[r,h,:] + ek[(r,c),h,:]) / 4 ;  softmax over the sources r of target c ;
//   out[c,h,:] = sum_r alpha * (v[r,h,:] + ev[(r,c),h,:])
// qkv rows are [head][q(16) | k(16) | v(16)] (lin_qkv viewed [H,3,C]), ekv rows [head][ek(16) | ev(16)].
template <typename AT, bool kFast, int MAXN>
__global__ void __launch_bounds__(256) k_wo_attention(Plan plan, int ngrp, const AT* __restrict__ qkv, const AT* __restrict__ ekv,
                                                      AT* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float sq[ATT_G][256];
  __shared__ float slog[ATT_G][MAXN][N_HEADS];
  const int mol = blockIdx.x / ngrp, j0 = (blockIdx.x % ngrp) * ATT_G;
  const int n = plan.n_atoms[mol];
  if (j0 >= n) return;
  const int t = threadIdx.x;
  const int base = plan.noff[mol];
  const size_t dbase = static_cast<size_t>(2) * plan.poff[mol];
  const int gsz = min(ATT_G, n - j0);
  for (int idx = t; idx < gsz * 64; idx += 256) {       // q of the targets, head-major [h][16]
    const int jl = idx >> 6, c4 = (idx & 63) * 4, hh = c4 >> 4, dd = c4 & 15;
    *reinterpret_cast<float4*>(&sq[jl][c4]) = load4<AT>(qkv + static_cast<size_t>(base + j0 + jl) * QKV_LD + hh * 48 + dd);
  }
  __syncthreads();
  const unsigned rcp_n = 65536u / static_cast<unsigned>(n) + 1u;
  for (int idx = t; idx < gsz * n * N_HEADS; idx += 256) {
    const int hh = idx & 15, r = idx >> 4;
    const int jl = static_cast<int>((static_cast<unsigned>(r) * rcp_n) >> 16), i = r - jl * n, j = j0 + jl;
    if (i == j) continue;
    const AT* kr = qkv + static_cast<size_t>(base + i) * QKV_LD + hh * 48 + 16;
    const AT* er = ekv + (dbase + static_cast<size_t>(i) * (n - 1) + (j - (j > i ? 1 : 0))) * E01_LD + hh * 32;
    const float* qr = &sq[jl][hh * 16];
    float a = 0.f;
    if constexpr (sizeof(AT) == 2) {
      // 16 channels = two 16-byte loads per operand, all four in flight; k + ek as packed bf16 adds, q and the sum fp32
      const uint4 k0 = *reinterpret_cast<const uint4*>(kr), k1 = *reinterpret_cast<const uint4*>(kr + 8);
      const uint4 e0 = *reinterpret_cast<const uint4*>(er), e1 = *reinterpret_cast<const uint4*>(er + 8);
      const __nv_bfloat162* kp0 = reinterpret_cast<const __nv_bfloat162*>(&k0);
      const __nv_bfloat162* kp1 = reinterpret_cast<const __nv_bfloat162*>(&k1);
      const __nv_bfloat162* ep0 = reinterpret_cast<const __nv_bfloat162*>(&e0);
      const __nv_bfloat162* ep1 = reinterpret_cast<const __nv_bfloat162*>(&e1);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float2 s0 = __bfloat1622float2(__hadd2(kp0[d], ep0[d])), s1 = __bfloat1622float2(__hadd2(kp1[d], ep1[d]));
        const float2 q0 = *reinterpret_cast<const float2*>(qr + 2 * d), q1 = *reinterpret_cast<const float2*>(qr + 8 + 2 * d);
        a = fmaf(q0.x, s0.x, a);
        a = fmaf(q0.y, s0.y, a);
        a = fmaf(q1.x, s1.x, a);
        a = fmaf(q1.y, s1.y, a);
      }
    } else {
#pragma unroll
      for (int d = 0; d < 16; d += 4) {
        const float4 kv = load4<AT>(kr + d), ev = load4<AT>(er + d);
        const float4 qv = *reinterpret_cast<const float4*>(qr + d);
        a = fmaf(qv.x, kv.x + ev.x, a);
        a = fmaf(qv.y, kv.y + ev.y, a);
        a = fmaf(qv.z, kv.z + ev.z, a);
        a = fmaf(qv.w, kv.w + ev.w, a);
      }
    }
    slog[jl][i][hh] = a * 0.25f;                         // 1 / sqrt(out_channels = 16)
  }
  __syncthreads();
  if (t < gsz * N_HEADS) {                               // softmax over sources: thread <-> (target, head)
    const int jl = t >> 4, hh = t & 15, j = j0 + jl;
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i)
      if (i != j) mx = fmaxf(mx, slog[jl][i][hh]);
    float den = 0.f;
    for (int i = 0; i < n; ++i) {
      const float ex = (i != j) ? act_exp<kFast>(slog[jl][i][hh] - mx) : 0.f;
      slog[jl][i][hh] = ex;
      den += ex;
    }
    const float inv = 1.0f / (den + 1e-16f);
    for (int i = 0; i < n; ++i) slog[jl][i][hh] *= inv;
  }
  __syncthreads();
  {                                                      // messages: warp <-> target, lane <-> 8 channels of one head
    const int w = t >> 5, lane = t & 31;
    if (w >= gsz) return;
    const int j = j0 + w, hh = lane >> 1, d0 = (lane & 1) * 8;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
      if (i == j) continue;
      const float al = slog[w][i][hh];
      const AT* vr = qkv + static_cast<size_t>(base + i) * QKV_LD + hh * 48 + 32 + d0;
      const AT* er = ekv + (dbase + static_cast<size_t>(i) * (n - 1) + (j - (j > i ? 1 : 0))) * E01_LD + hh * 32 + 16 + d0;
      if constexpr (sizeof(AT) == 2) {
        const uint4 vv = *reinterpret_cast<const uint4*>(vr), ee = *reinterpret_cast<const uint4*>(er);
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vv);
        const __nv_bfloat162* e2 = reinterpret_cast<const __nv_bfloat162*>(&ee);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 sv = __bfloat1622float2(__hadd2(v2[k], e2[k]));
          acc[2 * k] = fmaf(al, sv.x, acc[2 * k]);
          acc[2 * k + 1] = fmaf(al, sv.y, acc[2 * k + 1]);
        }
      } else {
        const float4 v0 = load4<AT>(vr), v1 = load4<AT>(vr + 4), e0 = load4<AT>(er), e1 = load4<AT>(er + 4);
        acc[0] = fmaf(al, v0.x + e0.x, acc[0]); acc[1] = fmaf(al, v0.y + e0.y, acc[1]);
        acc[2] = fmaf(al, v0.z + e0.z, acc[2]); acc[3] = fmaf(al, v0.w + e0.w, acc[3]);
        acc[4] = fmaf(al, v1.x + e1.x, acc[4]); acc[5] = fmaf(al, v1.y + e1.y, acc[5]);
        acc[6] = fmaf(al, v1.z + e1.z, acc[6]); acc[7] = fmaf(al, v1.w + e1.w, acc[7]);
      }
    }
    AT* o = out + static_cast<size_t>(base + j) * 256 + hh * 16 + d0;
    store4<AT>(o, acc[0], acc[1], acc[2], acc[3]);
    store4<AT>(o + 4, acc[4], acc[5], acc[6], acc[7]);
  }
}

// h1 = h_in + ng1 * h_node (kept un-normalised: it is the FFN's residual base) ; h1b = modulate(norm2_node(h1))   (:587-596)
template <typename AT, bool kFast>
__global__ void __launch_bounds__(256) k_wo_node_update1(Plan plan, const float* __restrict__ h, const float* __restrict__ hn,
                                                         const float* __restrict__ ada, int l, float* __restrict__ h1,
                                                         AT* __restrict__ h1b) {
  pdl_trigger();
  pdl_wait();
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= plan.Mn) return;
  const int mol = plan.node_info[m] >> 6;
  const float* ar = ada + static_cast<size_t>(mol) * ADA_LD + l * ADA_BLK + ADA_NODE;
  float v[8], a[8], g[8];
  load8(h + static_cast<size_t>(m) * 256, lane, v);
  load8(hn + static_cast<size_t>(m) * 256, lane, a);
  load8(ar + 512, lane, g);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = v[i] + g[i] * a[i];
  store8<float>(h1 + static_cast<size_t>(m) * 256, lane, v);
  ln256_mod<kFast>(v, ar + 768, ar + 1024, lane);
  store8<AT>(h1b + static_cast<size_t>(m) * 256, lane, v);
}

// e1 = e_in + eg1 * node2edge_lin(cat[hn_r, hn_c]) ; e1b = modulate(norm2_edge(e1))   (:603-622); half a warp per edge
template <typename AT, bool kFast>
__global__ void __launch_bounds__(256) k_wo_edge_update1(int Md, const int4* __restrict__ dir_info, const float* __restrict__ e,
                                                         const float* __restrict__ pab, const float* __restrict__ n2e_b,
                                                         const float* __restrict__ ada, int l, float* __restrict__ e1f,
                                                         AT* __restrict__ e1b) {
  pdl_trigger();
  pdl_wait();
  const int d = blockIdx.x * 16 + (threadIdx.x >> 4), c0 = (threadIdx.x & 15) * 4;
  const bool ok = d < Md;
  const int dd = ok ? d : Md - 1;
  const int4 info = dir_info[dd];
  const float* ar = ada + static_cast<size_t>(info.w) * ADA_LD + l * ADA_BLK + ADA_EDGE;
  const float4 ev = *reinterpret_cast<const float4*>(e + static_cast<size_t>(dd) * 64 + c0);
  const float4 pr = *reinterpret_cast<const float4*>(pab + static_cast<size_t>(info.y) * 128 + c0);
  const float4 pc = *reinterpret_cast<const float4*>(pab + static_cast<size_t>(info.z) * 128 + 64 + c0);
  const float4 bb = *reinterpret_cast<const float4*>(n2e_b + c0);
  const float4 g = *reinterpret_cast<const float4*>(ar + 128 + c0);
  const float4 sh = *reinterpret_cast<const float4*>(ar + 192 + c0);
  const float4 sc = *reinterpret_cast<const float4*>(ar + 256 + c0);
  float v0 = ev.x + g.x * ((pr.x + pc.x) + bb.x);
  float v1 = ev.y + g.y * ((pr.y + pc.y) + bb.y);
  float v2 = ev.z + g.z * ((pr.z + pc.z) + bb.z);
  float v3 = ev.w + g.w * ((pr.w + pc.w) + bb.w);
  if (ok) *reinterpret_cast<float4*>(e1f + static_cast<size_t>(d) * 64 + c0) = make_float4(v0, v1, v2, v3);
  const float mean = half_warp_sum((v0 + v1) + (v2 + v3)) * (1.0f / 64.0f);
  v0 -= mean; v1 -= mean; v2 -= mean; v3 -= mean;
  const float var = half_warp_sum((v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3)) * (1.0f / 64.0f);
  const float is = inv_std<kFast>(var);
  v0 = (v0 * is) * (1.0f + sc.x) + sh.x;
  v1 = (v1 * is) * (1.0f + sc.y) + sh.y;
  v2 = (v2 * is) * (1.0f + sc.z) + sh.z;
  v3 = (v3 * is) * (1.0f + sc.w) + sh.w;
  if (ok) store4<AT>(e1b + static_cast<size_t>(d) * 64 + c0, v0, v1, v2, v3);
}

// out = resid + gate[mol] * f (+ activation-dtype copy); the split form of the RESGATE epilogue (validation mode)
template <typename AT>
__global__ void k_wo_resgate(int rows, int W, const uint32_t* __restrict__ info, int shift, const float* __restrict__ resid,
                             const float* __restrict__ f, const float* __restrict__ ada_l, int gate_off, float* __restrict__ out,
                             AT* __restrict__ outb) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = W / 4;
  if (idx >= rows * per) return;
  const int r = idx / per, c = (idx % per) * 4;
  const float4 a = *reinterpret_cast<const float4*>(resid + static_cast<size_t>(r) * W + c);
  const float4 b = *reinterpret_cast<const float4*>(f + static_cast<size_t>(r) * W + c);
  const float4 g = *reinterpret_cast<const float4*>(ada_l + static_cast<size_t>(info[r] >> shift) * ADA_LD + gate_off + c);
  const float4 v = make_float4(a.x + g.x * b.x, a.y + g.y * b.y, a.z + g.z * b.z, a.w + g.w * b.w);
  *reinterpret_cast<float4*>(out + static_cast<size_t>(r) * W + c) = v;
  store4<AT>(outb + static_cast<size_t>(r) * W + c, v.x, v.y, v.z, v.w);
}

// pos = pos_pred_mlp.2(tanh(pos_pred_mlp.0(atom_hids)))   (:709-717, :915); the tanh layer is a GEMM epilogue
template <typename AT>
__global__ void __launch_bounds__(256) k_wo_pos_out(int Mn, const AT* __restrict__ t1, const float* __restrict__ w,
                                                    float* __restrict__ pos) {
  pdl_trigger();
  pdl_wait();
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= Mn) return;
  float v[8];
  load8<AT>(t1 + static_cast<size_t>(m) * 256, lane, v);
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    float ww[8];
    load8(w + o * 256, lane, ww);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(v[k], ww[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) pos[static_cast<size_t>(m) * 3 + o] = acc;
  }
}

// edge_final = (E + E^T) / 2 on the pair layout   (:921-924)
__global__ void k_wo_sym(Plan plan, const float* __restrict__ pred_dir, float* __restrict__ pred_e) {
  pdl_trigger();
  pdl_wait();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= plan.Mp) return;
  int mol, i, j;
  unpack_pair(plan.pair_info[p], mol, i, j);
  const int n = plan.n_atoms[mol];
  const size_t dbase = static_cast<size_t>(2) * plan.poff[mol];
  const size_t d1 = dbase + static_cast<size_t>(i) * (n - 1) + (j - 1), d2 = dbase + static_cast<size_t>(j) * (n - 1) + i;
  const float2 a = *reinterpret_cast<const float2*>(pred_dir + d1 * 2), b = *reinterpret_cast<const float2*>(pred_dir + d2 * 2);
  *reinterpret_cast<float2*>(pred_e + static_cast<size_t>(p) * 2) = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
}

template <typename AT, bool kFast>
int denoise_wo_impl(DsContext* ctx, const PackedWeights& pw, const Plan& plan, const float* xs, const float* es,
                    const float* cond_x, const float* cond_e, const float* noise_level, StepRef sr, const float* ctx_emb,
                    float* pred_x, float* pred_e, DenoiseWs& w, cudaStream_t s) {
  const int B = plan.B, Mn = plan.Mn, Mp = plan.Mp, Md = 2 * plan.Mp;
  const int AD = kFast ? DT_BF16 : DT_F32;
  AT* X = reinterpret_cast<AT*>(w.X);
  AT* hb = reinterpret_cast<AT*>(w.hb);
  AT* ahid = reinterpret_cast<AT*>(w.ahid);
  AT* ehid = reinterpret_cast<AT*>(w.ehid);
  AT* eb = reinterpret_cast<AT*>(w.eb);

  // out(fp32) = resid + gate * (A W^T + bias) with an activation-dtype copy; resid == null: plain A W^T + bias
  auto resgate = [&](const void* A, int lda, const void* W, int ldw, const float* bias, int M, int N, int K,
                     const uint32_t* info, int shift, const float* ada_l, int gate_off, const float* resid, float* out,
                     void* outb, int ldob, float* scratch) -> int {
    if (M <= 0) return DS_OK;
    if (kFast) {
      GemmDesc g;
      g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.out = out; g.ldo = N; g.M = M; g.N = N; g.K = K;
      g.a_dtype = DT_BF16; g.out_dtype = DT_F32; g.mode = GEMM_RESGATE; g.row_info = info; g.info_shift = shift;
      g.ada = ada_l; g.off_a = gate_off; g.resid = resid; g.ldres = N; g.out2 = outb; g.ldo2 = ldob;
      return gemm_tc_launch(ctx, g, s);
    }
    if (resid == nullptr) {
      DS_TRY(linear(ctx, A, lda, W, ldw, bias, nullptr, 0, out, N, DT_F32, M, N, K, ACT_NONE, s));
      if (outb && outb != static_cast<void*>(out)) {
        ds_launch(k_wo_copy<AT>, dim3(cdiv(M * (N / 4), 256)), dim3(256), 0, s, M, N, out, reinterpret_cast<AT*>(outb), ldob,
                  static_cast<AT*>(nullptr), 0);
        LAUNCH_CHECK(ctx);
      }
      return DS_OK;
    }
    DS_TRY(linear(ctx, A, lda, W, ldw, bias, nullptr, 0, scratch, N, DT_F32, M, N, K, ACT_NONE, s));
    ds_launch(k_wo_resgate<AT>, dim3(cdiv(M * (N / 4), 256)), dim3(256), 0, s, M, N, info, shift, resid, scratch, ada_l, gate_off,
              out, reinterpret_cast<AT*>(outb));
    LAUNCH_CHECK(ctx);
    return DS_OK;
  };

  ds_launch(k_zero_flags, dim3(1), dim3(32), 0, s, w.flags);
  LAUNCH_CHECK(ctx);
  ds_launch(k_time_feat<AT>, dim3(B), dim3(256), 0, s, noise_level, sr, pw.tm_freq, pw.tm1_w, pw.tm1_b, reinterpret_cast<AT*>(w.tfeat));
  LAUNCH_CHECK(ctx);
  DS_TRY(linear(ctx, w.tfeat, D_TIME, pw.tm3_w, D_TIME, pw.tm3_b, ctx_emb, D_TIME, w.s_act, D_TIME, AD, B, D_TIME, D_TIME,
                ACT_SILU, s));
  DS_TRY(linear(ctx, w.s_act, D_TIME, pw.w_ada, D_TIME, pw.b_ada, nullptr, 0, w.ada, ADA_LD, DT_F32, B, ADA_LD, D_TIME,
                ACT_NONE, s));
  // root: NodeEmbed -> h, atom_hids[0]
  ds_launch(k_wo_root_nodes<AT>, dim3(Mn), dim3(256), 0, s, Mn, xs, cond_x, pw.wo_x_w, pw.wo_x_b, pw.wo_pos_w, pw.wo_pos_b,
            reinterpret_cast<AT*>(w.f1));
  LAUNCH_CHECK(ctx);
  DS_TRY(linear(ctx, w.f1, 512, pw.wo_mlp_w, 512, pw.wo_mlp_b, nullptr, 0, w.h, 256, DT_F32, Mn, 256, 512, ACT_NONE, s));
  ds_launch(k_wo_copy<AT>, dim3(cdiv(Mn * 64, 256)), dim3(256), 0, s, Mn, 256, w.h, hb, 256, ahid, 768);
  LAUNCH_CHECK(ctx);
  if (Mp > 0) {
    // root edge embedding per pair (the inputs are symmetric), expanded to directed edges
    ds_launch(k_root_pair_flags, dim3(cdiv(Mp, 256)), dim3(256), 0, s, plan, cond_x, cond_e, w.pflags, w.flags);
    LAUNCH_CHECK(ctx);
    if (kFast) {
      AT* xr = reinterpret_cast<AT*>(w.xr);
      ds_launch(k_root_operand<AT, kFast>, dim3(cdiv(Mp, 32)), dim3(256), 0, s, plan, es, cond_x, cond_e, w.ada, w.flags, pw.root_means, pw.root_stds, xr);
      LAUNCH_CHECK(ctx);
      DS_TRY(linear(ctx, xr, 128, pw.root_w, 128, pw.edge_emb_b, nullptr, 0, w.y1, 64, DT_F32, Mp, 64, 128, ACT_NONE, s));
    } else {
      ds_launch(k_root_pairs<AT, kFast>, dim3(cdiv(Mp, 32)), dim3(256), 0, s, plan, es, cond_x, cond_e, w.ada, w.flags, pw.root_means,
                pw.root_stds, pw.edge_emb_w, pw.edge_emb_b, w.y1, X, ehid);
      LAUNCH_CHECK(ctx);
    }
    ds_launch(k_wo_expand_root<AT>, dim3(cdiv(Md * 16, 256)), dim3(256), 0, s, Md, plan.dir_info, w.y1, w.e, ehid);
    LAUNCH_CHECK(ctx);
  }

  const int ngrp = (plan.N + ATT_G - 1) / ATT_G;
  for (int l = 0; l < N_LAYERS; ++l) {
    const BlockWeights& bw = pw.blk[l];
    const float* ada_l = w.ada + l * ADA_BLK;
    ds_launch(k_node_ln1<AT, kFast>, dim3(cdiv(Mn, 8)), dim3(256), 0, s, plan, w.h, w.ada, l, reinterpret_cast<AT*>(w.hh));
    LAUNCH_CHECK(ctx);
    DS_TRY(linear(ctx, w.hh, 256, bw.wqkv, 256, bw.bqkv, nullptr, 0, w.qkv, QKV_LD, AD, Mn, QKV_LD, 256, ACT_NONE, s));
    if (Mp > 0) {
      ds_launch(k_wo_dir_ln1<AT, kFast>, dim3(cdiv(Md, 8)), dim3(256), 0, s, Md, plan.dir_mol, w.e, w.ada, l, reinterpret_cast<AT*>(w.ea));
      LAUNCH_CHECK(ctx);
      DS_TRY(linear(ctx, w.ea, 64, bw.wkve, 64, nullptr, nullptr, 0, w.e01, E01_LD, AD, Md, E01_LD, 64, ACT_NONE, s));
    }
    if (plan.N <= 32)
      ds_launch(k_wo_attention<AT, kFast, 32>, dim3(B * ngrp), dim3(256), 0, s, plan, ngrp, reinterpret_cast<const AT*>(w.qkv),
                reinterpret_cast<const AT*>(w.e01), reinterpret_cast<AT*>(w.hnb));
    else
      ds_launch(k_wo_attention<AT, kFast, 64>, dim3(B * ngrp), dim3(256), 0, s, plan, ngrp, reinterpret_cast<const AT*>(w.qkv),
                reinterpret_cast<const AT*>(w.e01), reinterpret_cast<AT*>(w.hnb));
    LAUNCH_CHECK(ctx);
    // h_node = proj(attention): fp32 for the residual, activation dtype for the hoisted node2edge_lin
    DS_TRY(resgate(w.hnb, 256, bw.wproj, 256, bw.bproj, Mn, 256, 256, nullptr, 0, nullptr, 0, nullptr, w.hn, w.h1b, 256, nullptr));
    DS_TRY(linear(ctx, w.h1b, 256, bw.wn2e2, 256, nullptr, nullptr, 0, w.pab, 128, DT_F32, Mn, 128, 256, ACT_NONE, s));
    // node update
    ds_launch(k_wo_node_update1<AT, kFast>, dim3(cdiv(Mn, 8)), dim3(256), 0, s, plan, w.h, w.hn, w.ada, l, w.h1, reinterpret_cast<AT*>(w.h1b));
    LAUNCH_CHECK(ctx);
    DS_TRY(linear(ctx, w.h1b, 256, bw.ff1_w, 256, bw.ff1_b, nullptr, 0, w.f1, 512, AD, Mn, 512, 256, ACT_GELU, s));
    DS_TRY(resgate(w.f1, 512, bw.ff2_w, 512, bw.ff2_b, Mn, 256, 512, plan.node_info, 6, ada_l, ADA_NODE + 1280, w.h1, w.h, hb, 256, w.f2));
    if (Mp > 0) {
      // edge update
      ds_launch(k_wo_edge_update1<AT, kFast>, dim3(cdiv(Md, 16)), dim3(256), 0, s, Md, plan.dir_info, w.e, w.pab, bw.n2e_b, w.ada, l,
                w.e1f, reinterpret_cast<AT*>(w.e1b));
      LAUNCH_CHECK(ctx);
      DS_TRY(linear(ctx, w.e1b, 64, bw.ff3_w, 64, bw.ff3_b, nullptr, 0, w.f3, 128, AD, Md, 128, 64, ACT_GELU, s));
      DS_TRY(resgate(w.f3, 128, bw.ff4_w, 128, bw.ff4_b, Md, 64, 128, plan.dir_mol, 0, ada_l, ADA_EDGE + 320, w.e1f, w.e, eb, 64, w.y1));
      DS_TRY(linear(ctx, eb, 64, bw.edge_w, 64, bw.edge_b, nullptr, 0, ehid + 64 + 16 * l, 192, AD, Md, 16, 64, ACT_NONE, s));
    }
    DS_TRY(linear(ctx, hb, 256, bw.node_w, 256, bw.node_b, nullptr, 0, ahid + 256 + 64 * l, 768, AD, Mn, 64, 256, ACT_NONE, s));
  }

  // heads
  DS_TRY(linear(ctx, ahid, 768, pw.np0_w, 768, pw.np0_b, nullptr, 0, w.n1, 256, AD, Mn, 256, 768, ACT_SILU, s));
  DS_TRY(linear(ctx, w.n1, 256, pw.np2_w, 256, pw.np2_b, nullptr, 0, w.n2, 128, AD, Mn, 128, 256, ACT_SILU, s));
  ds_launch(k_node_head_out<AT>, dim3(cdiv(Mn, 8)), dim3(256), 0, s, plan, reinterpret_cast<const AT*>(w.n2), pw.np4_w, pw.np4_b, pred_x);
  LAUNCH_CHECK(ctx);
  DS_TRY(linear(ctx, ahid, 768, pw.wo_p0_w, 768, nullptr, nullptr, 0, w.hh, 256, AD, Mn, 256, 768, ACT_TANH, s));
  ds_launch(k_wo_pos_out<AT>, dim3(cdiv(Mn, 8)), dim3(256), 0, s, Mn, reinterpret_cast<const AT*>(w.hh), pw.wo_p2_w, w.pos);
  LAUNCH_CHECK(ctx);
  if (Mp > 0) {
    DS_TRY(linear(ctx, ehid, 192, pw.eh0_w, 192, pw.eh0_b, nullptr, 0, w.eh1, 128, AD, Md, 128, 192, ACT_SILU, s));
    if (kFast) {
      GemmDesc g;
      g.A = w.eh1; g.lda = 128; g.W = pw.eh2_bd; g.ldw = 128; g.bias = pw.eh2_b; g.M = Md; g.N = 64; g.K = 128;
      g.a_dtype = DT_BF16; g.mode = GEMM_EHEAD; g.wc2 = pw.eh4_wb; g.wdir = w.pred_dir;
      DS_TRY(gemm_tc_launch(ctx, g, s));
    } else {
      ds_launch(k_edge_head_out<AT, kFast>, dim3(cdiv(Md, 8)), dim3(256), 0, s, Md, reinterpret_cast<const AT*>(w.eh1), pw.eh2t_w,
                pw.eh2_b, pw.eh4_w, pw.eh4_b, w.pred_dir);
      LAUNCH_CHECK(ctx);
    }
    ds_launch(k_wo_sym, dim3(cdiv(Mp, 256)), dim3(256), 0, s, plan, w.pred_dir, pred_e);
    LAUNCH_CHECK(ctx);
  }
  ds_launch(k_pos_nan_flag, dim3(cdiv(Mn * 3, 256)), dim3(256), 0, s, Mn, w.pos, w.flags);
  LAUNCH_CHECK(ctx);
  ds_launch(k_pos_final, dim3(B), dim3(64), 0, s, plan, w.pos, w.flags, pred_x);
  LAUNCH_CHECK(ctx);
  return DS_OK;
}
