// Re-packing of a reference-compatible DMT state_dict (fp32 tensors on the device, addressed BY NAME) into
// the GEMM-ready blob used by the kernels.  Names/shapes = SURVEY.md Appendix B (models/dmt.py:211-262,
// models/layers.py:115-120, models/specformer.py:139-147,314-333,444-446).
#include <string.h>

#include "kernels.cuh"

namespace {

template <typename TD>
__global__ void k_copy2d(const float* __restrict__ src, int src_ld, TD* __restrict__ dst, int dst_ld, int rows, int cols,
                         float scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx % cols;
  dst[static_cast<size_t>(r) * dst_ld + c] = from_f32<TD>(scale * src[static_cast<size_t>(r) * src_ld + c]);
}
// dst[c, r] = src[r, c]
__global__ void k_transpose(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx % cols;
  dst[static_cast<size_t>(c) * rows + r] = src[idx];
}

// dst row head_perm(r) <- src row r  (matrix [QK_DIM, cols]);  vec: dst[head_perm(c)] <- src[c]
template <typename T>
__global__ void k_copy_headperm(const float* __restrict__ src, int src_ld, T* __restrict__ dst, int dst_ld, int rows, int cols,
                                int vec) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx % cols;
  const float v = src[static_cast<size_t>(r) * src_ld + c];
  if (vec)
    dst[head_perm(c)] = from_f32<T>(v);
  else
    dst[static_cast<size_t>(head_perm(r)) * dst_ld + c] = from_f32<T>(v);
}

struct Packer {
  DsContext* ctx;
  std::map<std::string, const float*> params;
  Arena arena;
  cudaStream_t s;
  bool bf;
  int err = DS_OK;

  const float* get(const std::string& name) {
    auto it = params.find(name);
    if (it == params.end()) {
      if (err == DS_OK) {
        ds_set_error("ds_pack_weights: parameter '%s' missing from the state_dict", name.c_str());
        err = DS_ERR_MISSING_PARAM;
      }
      return nullptr;
    }
    return it->second;
  }
  size_t es() const { return bf ? 2 : 4; }
  void* alloc_act(size_t elems) { return arena.take(elems * es()); }
  float* alloc_f32(size_t elems) { return static_cast<float*>(arena.take(elems * 4)); }

  // copy a [rows, cols] sub-block (src pre-offset, leading dim src_ld) into dst (+ element offset), dtype act or f32
  void copy(const float* src, int src_ld, void* dst, size_t dst_off, int dst_ld, int rows, int cols, bool act,
            float scale = 1.0f) {
    if (arena.dry || err != DS_OK || src == nullptr || dst == nullptr) return;
    const int n = rows * cols;
    if (act && bf)
      k_copy2d<bf16><<<(n + 255) / 256, 256, 0, s>>>(src, src_ld, static_cast<bf16*>(dst) + dst_off, dst_ld, rows, cols, scale);
    else
      k_copy2d<float><<<(n + 255) / 256, 256, 0, s>>>(src, src_ld, static_cast<float*>(dst) + dst_off, dst_ld, rows, cols, scale);
  }
  // the head-interleaved order of q / k / e0 (common.cuh: head_perm): [QK_DIM, cols] weight rows, or a QK_DIM bias vector
  void copy_headperm(const float* src, int src_ld, void* dst, size_t dst_off, int dst_ld, int cols, bool act, bool vec) {
    if (arena.dry || err != DS_OK || src == nullptr || dst == nullptr) return;
    const int rows = vec ? 1 : QK_DIM, cc = vec ? QK_DIM : cols, n = rows * cc;
    if (act && bf)
      k_copy_headperm<bf16><<<(n + 255) / 256, 256, 0, s>>>(src, src_ld, static_cast<bf16*>(dst) + dst_off, dst_ld, rows, cc, vec ? 1 : 0);
    else
      k_copy_headperm<float><<<(n + 255) / 256, 256, 0, s>>>(src, src_ld, static_cast<float*>(dst) + dst_off, dst_ld, rows, cc, vec ? 1 : 0);
  }
  // whole matrix [rows, cols] in the act dtype
  const void* mat(const std::string& name, int rows, int cols, float scale = 1.0f) {
    void* d = alloc_act(static_cast<size_t>(rows) * cols);
    copy(get(name), cols, d, 0, cols, rows, cols, true, scale);
    return d;
  }
  const float* vec(const std::string& name, int n, float scale = 1.0f) {
    float* d = alloc_f32(n);
    copy(get(name), n, d, 0, n, 1, n, false, scale);
    return d;
  }
};

int build(Packer& P, PackedWeights& pw) {
  DsContext* ctx = P.ctx;
  const bool wo = ctx->model_kind == 1;      // DMT_WO_EQ parameter names (models/dmt_wo_eq.py:679-746)
  if (!wo) {
    pw.node_emb_w = P.vec("node_emb.weight", 256 * 12);
    pw.node_emb_b = P.vec("node_emb.bias", 256);
  } else {
    pw.wo_x_w = P.vec("node_emb.x_linear.weight", 512 * 12);
    pw.wo_x_b = P.vec("node_emb.x_linear.bias", 512);
    pw.wo_pos_w = P.vec("node_emb.pos_linear.weight", 512 * 3);
    pw.wo_pos_b = P.vec("node_emb.pos_linear.bias", 512);
    pw.wo_mlp_w = P.mat("node_emb.mlp.1.weight", 256, 512);
    pw.wo_mlp_b = P.vec("node_emb.mlp.1.bias", 256);
    pw.wo_p0_w = P.mat("pos_pred_mlp.0.weight", 256, 768);
    pw.wo_p2_w = P.vec("pos_pred_mlp.2.weight", 3 * 256);
  }
  pw.edge_emb_w = P.vec("edge_emb.weight", 64 * 68);
  pw.edge_emb_b = P.vec("edge_emb.bias", 64);
  {   // tensor-core form of the root edge embedding: operand columns [d0(64) | edge_x(2) cond_edge(2) | 0...]
    const float* we = P.get("edge_emb.weight");   // [64,68], input order [edge_x(2) | cond_edge(2) | d0(64)]
    void* wr = P.alloc_act(64 * 128);
    P.copy(we ? we + 4 : nullptr, 68, wr, 0, 128, 64, 64, true);
    P.copy(we, 68, wr, 64, 128, 64, 4, true);
    pw.root_w = wr;
  }
  pw.root_means = P.vec("dist_layer.means.weight", 63);
  pw.root_stds = P.vec("dist_layer.stds.weight", 63);
  pw.tm_freq = P.vec("time_mlp.0.weights", 8);
  pw.tm1_w = P.vec("time_mlp.1.weight", 1024 * 17);
  pw.tm1_b = P.vec("time_mlp.1.bias", 1024);
  pw.tm3_w = P.mat("time_mlp.3.weight", 1024, 1024);
  pw.tm3_b = P.vec("time_mlp.3.bias", 1024);

  // per-molecule adaLN table projection: all SiLU->Linear(1024, .) heads of the model stacked row-wise
  void* wada = P.alloc_act(static_cast<size_t>(ADA_LD) * D_TIME);
  float* bada = P.alloc_f32(ADA_LD);
  pw.w_ada = wada;
  pw.b_ada = bada;
  auto ada_rows = [&](const std::string& prefix, int row0, int rows) {
    P.copy(P.get(prefix + ".weight"), D_TIME, wada, static_cast<size_t>(row0) * D_TIME, D_TIME, rows, D_TIME, true);
    P.copy(P.get(prefix + ".bias"), rows, bada, row0, rows, 1, rows, false);
  };
  ada_rows("dist_layer.time_mlp.1", ADA_ROOT_RBF, 2);

  char buf[128];
  for (int l = 0; wo && l < N_LAYERS; ++l) {
    BlockWeights& b = pw.blk[l];
    snprintf(buf, sizeof(buf), "dmt_block_%d.", l);
    const std::string p(buf);
    ada_rows(p + "node_time_mlp.1", l * ADA_BLK + ADA_NODE, 1536);
    ada_rows(p + "edge_time_mlp.1", l * ADA_BLK + ADA_EDGE, 384);
    b.wqkv = P.mat(p + "attn_mpnn.lin_qkv.weight", 768, 256);
    b.bqkv = P.vec(p + "attn_mpnn.lin_qkv.bias", 768);
    b.wkve = P.mat(p + "attn_mpnn.lin_kv_e.weight", 512, 64);
    b.wproj = P.mat(p + "attn_mpnn.proj.weight", 256, 256);
    b.bproj = P.vec(p + "attn_mpnn.proj.bias", 256);
    {   // node2edge_lin [64,512] acts on cat[hn_r, hn_c]: applied per atom as two 256 -> 64 projections
      const float* w = P.get(p + "node2edge_lin.weight");
      void* w2 = P.alloc_act(128 * 256);
      P.copy(w, 512, w2, 0, 256, 64, 256, true);
      P.copy(w ? w + 256 : nullptr, 512, w2, 64 * 256, 256, 64, 256, true);
      b.wn2e2 = w2;
      b.n2e_b = P.vec(p + "node2edge_lin.bias", 64);
    }
    b.ff1_w = P.mat(p + "ff_linear1.weight", 512, 256);
    b.ff1_b = P.vec(p + "ff_linear1.bias", 512);
    b.ff2_w = P.mat(p + "ff_linear2.weight", 256, 512);
    b.ff2_b = P.vec(p + "ff_linear2.bias", 256);
    b.ff3_w = P.mat(p + "ff_linear3.weight", 128, 64);
    b.ff3_b = P.vec(p + "ff_linear3.bias", 128);
    b.ff4_w = P.mat(p + "ff_linear4.weight", 64, 128);
    b.ff4_b = P.vec(p + "ff_linear4.bias", 64);
    snprintf(buf, sizeof(buf), "node_%d", l);
    b.node_w = P.mat(std::string(buf) + ".weight", 64, 256);
    b.node_b = P.vec(std::string(buf) + ".bias", 64);
    snprintf(buf, sizeof(buf), "edge_%d", l);
    b.edge_w = P.mat(std::string(buf) + ".weight", 16, 64);
    b.edge_b = P.vec(std::string(buf) + ".bias", 16);
  }
  for (int l = 0; !wo && l < N_LAYERS; ++l) {
    BlockWeights& b = pw.blk[l];
    snprintf(buf, sizeof(buf), "e_block_%d.", l);
    const std::string p(buf);
    ada_rows(p + "node_time_mlp.1", l * ADA_BLK + ADA_NODE, 1536);
    ada_rows(p + "edge_time_mlp.1", l * ADA_BLK + ADA_EDGE, 384);
    ada_rows(p + "equi_update.time_mlp.1", l * ADA_BLK + ADA_COORD, 512);
    ada_rows(p + "dist_layer.time_mlp.1", l * ADA_BLK + ADA_RBF, 2);

    b.edge_emb_w = P.mat(p + "edge_emb.weight", 64, 128);
    b.edge_emb_b = P.vec(p + "edge_emb.bias", 64);
    void* w01 = P.alloc_act(static_cast<size_t>(E01_LD) * 64);
    P.copy_headperm(P.get(p + "attn_mpnn.lin_edge0.weight"), 64, w01, 0, 64, 64, true, false);
    P.copy(P.get(p + "attn_mpnn.lin_edge1.weight"), 64, w01, 256 * 64, 64, 256, 64, true);
    b.w01 = w01;
    void* wqkv = P.alloc_act(static_cast<size_t>(QKV_LD) * 256);
    float* bqkv = P.alloc_f32(QKV_LD);
    P.copy_headperm(P.get(p + "attn_mpnn.lin_query.weight"), 256, wqkv, 0, 256, 256, true, false);
    P.copy_headperm(P.get(p + "attn_mpnn.lin_key.weight"), 256, wqkv, 256 * 256, 256, 256, true, false);
    P.copy(P.get(p + "attn_mpnn.lin_value.weight"), 256, wqkv, 512 * 256, 256, 256, 256, true);
    P.copy_headperm(P.get(p + "attn_mpnn.lin_query.bias"), QK_DIM, bqkv, 0, QK_DIM, QK_DIM, false, true);
    P.copy_headperm(P.get(p + "attn_mpnn.lin_key.bias"), QK_DIM, bqkv, 256, QK_DIM, QK_DIM, false, true);
    P.copy(P.get(p + "attn_mpnn.lin_value.bias"), 256, bqkv, 512, 256, 1, 256, false);
    b.wqkv = wqkv;
    b.bqkv = bqkv;
    b.n2e_w = P.mat(p + "node2edge_lin.weight", 64, 256);
    b.n2e_b = P.vec(p + "node2edge_lin.bias", 64);
    b.ff1_w = P.mat(p + "ff_linear1.weight", 512, 256);
    b.ff1_b = P.vec(p + "ff_linear1.bias", 512);
    b.ff2_w = P.mat(p + "ff_linear2.weight", 256, 512);
    b.ff2_b = P.vec(p + "ff_linear2.bias", 256);
    // bf16 mode: ff_linear3 is stored halved like coord_mlp.0 (its SiLU runs as h + h tanh(h), ACT_SILU_HALF)
    b.ff3_w = P.mat(p + "ff_linear3.weight", 128, 64, P.bf ? 0.5f : 1.0f);
    b.ff3_b = P.vec(p + "ff_linear3.bias", 128, P.bf ? 0.5f : 1.0f);
    b.ff4_w = P.mat(p + "ff_linear4.weight", 64, 128);
    b.ff4_b = P.vec(p + "ff_linear4.bias", 64);
    // equi_update.input_lin [256, 640]: columns = [h_row(256) | h_col(256) | e(64) | dist(64)]
    const float* wil = P.get(p + "equi_update.input_lin.weight");
    void* we = P.alloc_act(256 * 128);
    P.copy(wil ? wil + 576 : nullptr, 640, we, 0, 128, 256, 64, true);    // dist columns first
    P.copy(wil ? wil + 512 : nullptr, 640, we, 64, 128, 256, 64, true);   // then e columns
    b.we = we;
    // rows 0..511: h_row | h_col parts of input_lin; rows 512..575: the block's skip projection node_l (dmt.py:387-388), so
    // that ONE GEMM over the updated atom rows produces both (the skip columns go to the atom-head operand)
    void* wab = P.alloc_act(576 * 256);
    float* bab = P.alloc_f32(576);
    P.copy(wil, 640, wab, 0, 256, 256, 256, true);
    P.copy(wil ? wil + 256 : nullptr, 640, wab, 256 * 256, 256, 256, 256, true);
    P.copy(P.get(p + "equi_update.input_lin.bias"), 256, bab, 0, 256, 1, 256, false);
    snprintf(buf, sizeof(buf), "node_%d", l);
    P.copy(P.get(std::string(buf) + ".weight"), 256, wab, 512 * 256, 256, 64, 256, true);
    P.copy(P.get(std::string(buf) + ".bias"), 64, bab, 512, 64, 1, 64, false);
    b.wab = wab;
    b.bab = bab;
    // bf16 mode: coord_mlp.0 is stored halved (exact in bf16) so that its SiLU is h + h tanh(h) (ACT_SILU_HALF)
    b.wc1 = P.mat(p + "equi_update.coord_mlp.0.weight", 256, 256, P.bf ? 0.5f : 1.0f);
    b.bc1 = P.vec(p + "equi_update.coord_mlp.0.bias", 256, P.bf ? 0.5f : 1.0f);
    b.wc2 = P.vec(p + "equi_update.coord_mlp.2.weight", 3 * 256);
    b.coord_scale = P.vec(p + "equi_update.coord_norm.scale", 1);
    b.rbf_means = P.vec(p + "dist_layer.means.weight", 63);
    b.rbf_stds = P.vec(p + "dist_layer.stds.weight", 63);
    snprintf(buf, sizeof(buf), "node_%d", l);
    b.node_w = P.mat(std::string(buf) + ".weight", 64, 256);
    b.node_b = P.vec(std::string(buf) + ".bias", 64);
    snprintf(buf, sizeof(buf), "edge_%d", l);
    b.edge_w = P.mat(std::string(buf) + ".weight", 16, 64);
    b.edge_b = P.vec(std::string(buf) + ".bias", 16);
  }

  pw.np0_w = P.mat("node_pred_mlp.0.weight", 256, 768);
  pw.np0_b = P.vec("node_pred_mlp.0.bias", 256);
  pw.np2_w = P.mat("node_pred_mlp.2.weight", 128, 256);
  pw.np2_b = P.vec("node_pred_mlp.2.bias", 128);
  pw.np4_w = P.vec("node_pred_mlp.4.weight", 6 * 128);
  pw.np4_b = P.vec("node_pred_mlp.4.bias", 6);
  {
    void* w = P.alloc_act(128 * 192);
    float* b = P.alloc_f32(128);
    P.copy(P.get("edge_exist_mlp.0.weight"), 192, w, 0, 192, 64, 192, true);
    P.copy(P.get("edge_type_mlp.0.weight"), 192, w, 64 * 192, 192, 64, 192, true);
    P.copy(P.get("edge_exist_mlp.0.bias"), 64, b, 0, 64, 1, 64, false);
    P.copy(P.get("edge_type_mlp.0.bias"), 64, b, 64, 64, 1, 64, false);
    pw.eh0_w = w;
    pw.eh0_b = b;
    float* w2t = P.alloc_f32(2 * 64 * 32);
    float* b2 = P.alloc_f32(64);
    float* w4 = P.alloc_f32(64);
    float* b4 = P.alloc_f32(2);
    const char* heads[2] = {"edge_exist_mlp", "edge_type_mlp"};
    for (int hd = 0; hd < 2; ++hd) {
      const std::string h(heads[hd]);
      const float* w2 = P.get(h + ".2.weight");   // [32,64]
      if (!P.arena.dry && P.err == DS_OK && w2 && w2t)
        k_transpose<<<(32 * 64 + 255) / 256, 256, 0, P.s>>>(w2, w2t + hd * 64 * 32, 32, 64);
      P.copy(P.get(h + ".2.bias"), 32, b2, hd * 32, 32, 1, 32, false);
      P.copy(P.get(h + ".4.weight"), 32, w4, hd * 32, 32, 1, 32, false);
      P.copy(P.get(h + ".4.bias"), 1, b4, hd, 1, 1, 1, false);
    }
    pw.eh2t_w = w2t;
    pw.eh2_b = b2;
    pw.eh4_w = w4;
    pw.eh4_b = b4;
    // tensor-core form of the same layers: block-diagonal [64,128] second layer, [w4_exist | w4_type | b4] vector
    void* wbd = P.alloc_act(64 * 128);
    float* w4b = P.alloc_f32(66);
    P.copy(P.get("edge_exist_mlp.2.weight"), 64, wbd, 0, 128, 32, 64, true);
    P.copy(P.get("edge_type_mlp.2.weight"), 64, wbd, 32 * 128 + 64, 128, 32, 64, true);
    P.copy(P.get("edge_exist_mlp.4.weight"), 32, w4b, 0, 32, 1, 32, false);
    P.copy(P.get("edge_type_mlp.4.weight"), 32, w4b, 32, 32, 1, 32, false);
    P.copy(P.get("edge_exist_mlp.4.bias"), 1, w4b, 64, 1, 1, 1, false);
    P.copy(P.get("edge_type_mlp.4.bias"), 1, w4b, 65, 1, 1, 1, false);
    pw.eh2_bd = wbd;
    pw.eh4_wb = w4b;
  }

  // ---------------- SpecFormer (models/specformer.py) ----------------
  static const int kLen[3] = {701, 3501, 3501}, kPatch[3] = {20, 50, 50}, kStride[3] = {10, 25, 25};
  static const char* kPosName[3] = {"W_pos_uv", "W_pos_ir", "W_pos_raman"};
  const int v = ctx->spectra_version;
  pw.n_spec = (v == 3) ? 3 : 1;
  pw.q_len = 0;
  for (int sidx = 0; sidx < pw.n_spec; ++sidx) {
    const int ty = (v == 3) ? sidx : v;
    pw.spec_type[sidx] = ty;
    pw.patch_num[sidx] = (kLen[ty] - kPatch[ty]) / kStride[ty] + 1;
    pw.q_len += pw.patch_num[sidx];
    snprintf(buf, sizeof(buf), "cond_encoder.backbone.W_P.%d.", sidx);
    pw.wp_w[sidx] = P.vec(std::string(buf) + "weight", 128 * kPatch[ty]);
    pw.wp_b[sidx] = P.vec(std::string(buf) + "bias", 128);
    const std::string pos_name = std::string("cond_encoder.backbone.") + ((v == 3) ? kPosName[ty] : "W_pos");
    pw.w_pos[sidx] = P.vec(pos_name, pw.patch_num[sidx] * 128);
  }
  for (int l = 0; l < 3; ++l) {
    SpecLayerWeights& sl = pw.sl[l];
    snprintf(buf, sizeof(buf), "cond_encoder.backbone.encoder.layers.%d.", l);
    const std::string p(buf);
    void* wqkv = P.alloc_act(384 * 128);
    float* bqkv = P.alloc_f32(384);
    const char* qkvn[3] = {"W_Q", "W_K", "W_V"};
    for (int i = 0; i < 3; ++i) {
      P.copy(P.get(p + "self_attn." + qkvn[i] + ".weight"), 128, wqkv, static_cast<size_t>(i) * 128 * 128, 128, 128, 128, true);
      P.copy(P.get(p + "self_attn." + qkvn[i] + ".bias"), 128, bqkv, i * 128, 128, 1, 128, false);
    }
    sl.wqkv = wqkv;
    sl.bqkv = bqkv;
    sl.scale = P.vec(p + "self_attn.sdp_attn.scale", 1);
    sl.wo = P.mat(p + "self_attn.to_out.0.weight", 128, 128);
    sl.bo = P.vec(p + "self_attn.to_out.0.bias", 128);
    const char* bnn[2] = {"norm_attn.1.", "norm_ffn.1."};
    for (int i = 0; i < 2; ++i) {
      float* bn = P.alloc_f32(4 * 128);
      P.copy(P.get(p + bnn[i] + "weight"), 128, bn, 0, 128, 1, 128, false);
      P.copy(P.get(p + bnn[i] + "bias"), 128, bn, 128, 128, 1, 128, false);
      P.copy(P.get(p + bnn[i] + "running_mean"), 128, bn, 256, 128, 1, 128, false);
      P.copy(P.get(p + bnn[i] + "running_var"), 128, bn, 384, 128, 1, 128, false);
      (i == 0 ? sl.bn1 : sl.bn2) = bn;
    }
    sl.wf0 = P.mat(p + "ff.0.weight", 256, 128);
    sl.bf0 = P.vec(p + "ff.0.bias", 256);
    sl.wf3 = P.mat(p + "ff.3.weight", 128, 256);
    sl.bf3 = P.vec(p + "ff.3.bias", 128);
  }
  pw.head_w = P.mat("cond_encoder.head.linear.weight", 256, pw.q_len * 128);
  pw.head_b = P.vec("cond_encoder.head.linear.bias", 256);
  {
    float* on = P.alloc_f32(512);
    P.copy(P.get("cond_encoder.out_norm.weight"), 256, on, 0, 256, 1, 256, false);
    P.copy(P.get("cond_encoder.out_norm.bias"), 256, on, 256, 256, 1, 256, false);
    pw.out_norm = on;
  }
  pw.cond_w = P.mat("cond_lin.weight", 1024, 256);
  pw.cond_b = P.vec("cond_lin.bias", 1024);
  return P.err;
}

}  // namespace

size_t packed_weights_bytes(DsContext* ctx) {
  Packer P;
  P.ctx = ctx;
  P.arena = Arena{nullptr, 0, 0, true};
  P.s = nullptr;
  P.bf = ds_is_bf16(ctx);
  PackedWeights tmp;
  build(P, tmp);
  return P.arena.off + 256;
}

int pack_weights(DsContext* ctx, const char* const* names, const void* const* ptrs, int n, void* blob, size_t blob_bytes,
                 PackedWeights* out, cudaStream_t s) {
  DS_CHECK(blob != nullptr && (reinterpret_cast<uintptr_t>(blob) & 255) == 0, DS_ERR_INVALID,
           "ds_pack_weights: blob must be a 256-byte aligned device buffer");
  const size_t need = packed_weights_bytes(ctx);
  DS_CHECK(blob_bytes >= need, DS_ERR_WORKSPACE, "ds_pack_weights: blob too small (%zu < %zu)", blob_bytes, need);
  Packer P;
  P.ctx = ctx;
  P.arena = Arena{static_cast<uint8_t*>(blob), 0, blob_bytes, false};
  P.s = s;
  P.bf = ds_is_bf16(ctx);
  for (int i = 0; i < n; ++i) {
    const char* nm = names[i];
    if (strncmp(nm, "module.", 7) == 0) nm += 7;    // DataParallel-wrapped checkpoints (utils.py:15-19)
    P.params[nm] = static_cast<const float*>(ptrs[i]);
  }
  DS_CUDA_CHECK(cudaMemsetAsync(blob, 0, need, s));   // zero rows of the padded matrices
  PackedWeights pw;
  int r = build(P, pw);
  if (r != DS_OK) return r;
  DS_CUDA_CHECK(cudaGetLastError());
  pw.valid = true;
  *out = pw;
  return DS_OK;
}
