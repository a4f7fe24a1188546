// Internal interface: packed weights, workspace carving and kernel launchers shared by the API layer.
#pragma once
#include "context.cuh"

// ----------------------------------------------------------------------------- packed weights
// "act" pointers hold GEMM B operands in the activation dtype of the mode (bf16 in production mode,
// fp32 in validation mode), row-major [N, K] (= nn.Linear.weight layout, K-major for UMMA).
struct BlockWeights {
  const void* edge_emb_w;   // [64,128]  input cols = [dist(64) | e(64)]      dmt.py:80,139
  const float* edge_emb_b;
  const void* w01;          // [512,64]  rows 0..251 lin_edge0, 256..511 lin_edge1  layers.py:119-120
  const void* wqkv;         // [768,256] rows 0..251 query, 256..507 key, 512..767 value  layers.py:115-117
  const float* bqkv;        // [768]
  const void* n2e_w;        // [64,256]  node2edge_lin (applied per node, hoisted)   dmt.py:81,156-157
  const float* n2e_b;
  const void* ff1_w; const float* ff1_b;   // [512,256]
  const void* ff2_w; const float* ff2_b;   // [256,512]
  const void* ff3_w; const float* ff3_b;   // [128,64]
  const void* ff4_w; const float* ff4_b;   // [64,128]
  const void* we;           // [256,128] input_lin columns for [dist | e] (re-ordered)   dmt.py:27,39
  const void* wab;          // [512,256] rows 0..255 = input_lin[:, 0:256] (h_row), 256..511 = input_lin[:, 256:512] (h_col)
  const float* bab;         // [512]     input_lin.bias in the first half, 0 in the second
  const void* wc1; const float* bc1;       // coord_mlp.0 [256,256]
  const float* wc2;         // coord_mlp.2 [3,256] fp32
  const float* coord_scale; // [1]
  const float* rbf_means;   // [63]
  const float* rbf_stds;    // [63]
  const void* node_w; const float* node_b;   // node_l [64,256]
  const void* edge_w; const float* edge_b;   // edge_l [16,64]
  // DMT_WO_EQ (models/dmt_wo_eq.py:200-202,436-441): wqkv = lin_qkv [768,256] in the reference's [head][q|k|v][16]
  // order, n2e_b = node2edge_lin.bias
  const void* wkve;         // lin_kv_e [512,64], [head][k|v][16]
  const void* wproj; const float* bproj;   // proj [256,256]
  const void* wn2e2;        // [128,256] rows 0..63 = node2edge_lin[:, 0:256] (source), 64..127 = [:, 256:512] (target)
};

struct SpecLayerWeights {
  const void* wqkv; const float* bqkv;   // [384,128]
  const float* scale;                    // sdp_attn.scale (scalar)
  const void* wo; const float* bo;       // [128,128]
  const float* bn1;                      // [4,128] = weight, bias, running_mean, running_var
  const void* wf0; const float* bf0;     // [256,128]
  const void* wf3; const float* bf3;     // [128,256]
  const float* bn2;
};

struct PackedWeights {
  bool valid = false;
  // root
  const float* node_emb_w; const float* node_emb_b;   // [256,12]
  const float* edge_emb_w; const float* edge_emb_b;   // [64,68]
  const float* root_means; const float* root_stds;
  const void* w_ada; const float* b_ada;              // [ADA_LD,1024]
  const float* tm_freq;                               // [8]
  const float* tm1_w; const float* tm1_b;             // [1024,17]
  const void* tm3_w; const float* tm3_b;              // [1024,1024]
  BlockWeights blk[N_LAYERS];
  // heads
  const void* np0_w; const float* np0_b;   // [256,768]
  const void* np2_w; const float* np2_b;   // [128,256]
  const float* np4_w; const float* np4_b;  // [6,128]
  const void* eh0_w; const float* eh0_b;   // [128,192] rows 0..63 exist, 64..127 type
  const float* eh2t_w;                     // [2][64][32] transposed second layers (exist, type)
  const float* eh2_b;                      // [2][32]
  const float* eh4_w;                      // [2][32]
  const float* eh4_b;                      // [2]
  const void* eh2_bd;                      // act [64,128] block-diagonal second layers (tensor-core path)
  const float* eh4_wb;                     // [66] = w4 exist | w4 type | b4 exist, b4 type
  const void* root_w;                      // act [64,128] root edge_emb for operand [d0 | edge_x cond_edge | 0]
  // DMT_WO_EQ root / position head (models/dmt_wo_eq.py:629-643,709-717)
  const float* wo_x_w; const float* wo_x_b;      // node_emb.x_linear [512,12]
  const float* wo_pos_w; const float* wo_pos_b;  // node_emb.pos_linear [512,3]
  const void* wo_mlp_w; const float* wo_mlp_b;   // node_emb.mlp.1 [256,512]
  const void* wo_p0_w;                           // pos_pred_mlp.0 [256,768] (no bias)
  const float* wo_p2_w;                          // pos_pred_mlp.2 [3,256] (no bias)
  // SpecFormer
  int n_spec = 0;                          // number of spectra used (1 or 3)
  int spec_type[3];                        // 0 uv, 1 ir, 2 raman
  int patch_num[3];
  int q_len = 0;
  const float* wp_w[3]; const float* wp_b[3];   // [128, patch_len]
  const float* w_pos[3];                        // [patch_num,128]
  SpecLayerWeights sl[3];
  const void* head_w; const float* head_b;      // [256, q_len*128]
  const float* out_norm;                        // [2,256] weight, bias
  const void* cond_w; const float* cond_b;      // [1024,256]
};

// ----------------------------------------------------------------------------- workspace
struct Arena {
  uint8_t* base;
  size_t off, cap;
  bool dry;   // dry run: only measure
  void* take(size_t bytes) {
    size_t o = (off + 255) & ~size_t(255);
    off = o + bytes;
    if (dry || off > cap) return nullptr;
    return base + o;
  }
};

struct DenoiseWs {      // scratch of one denoiser call on a plan (sizes in elements of the given type)
  float* tfeat_f;       // unused in bf16
  void* tfeat;          // act [B,1024]
  void* s_act;          // act [B,1024]
  float* ada;           // [B,ADA_LD]
  float* h;             // [Mn,256] residual stream
  void* hb;             // act [Mn,256]
  float* h1;            // [Mn,256]
  void* h1b;            // act
  float* pos;           // [Mn,3]
  void* hh;             // act [Mn,256]
  void* qkv;            // [Mn,768] fp32 (validation) / bf16
  float* hn;            // [Mn,256]
  void* hnb;            // act
  float* pn;            // [Mn,64]
  void* f1;             // act [Mn,512]
  float* f2;            // [Mn,256]
  void* ab;             // [Mn,512] fp32 (validation) / bf16
  void* ahid;           // act [Mn,768]
  void* n1;             // act [Mn,256]
  void* n2;             // act [Mn,128]
  float* e;             // [Mp,64] residual stream
  float* e1f;           // [Mp,64]
  void* e1b;            // act [Mp,64]
  void* X;              // act [Mp,128] = [dist | e]
  float* y1;            // [Mp,64]  (edge_emb out, later ff4 out)
  void* ea;             // act [Mp,64]
  void* e01;            // act [Mp,512]
  void* f3;             // act [Mp,128]
  void* gp;             // act [Mp,256]
  void* ehid;           // act [Mp,192]
  void* eh1;            // act [Mp,128]
  void* xr;             // act [Mp,128] root edge-embedding operand
  uint8_t* pflags;      // [Mp]
  void* Z;              // act [2Mp,256]
  void* u1;             // act [2Mp,256]
  float* wdir;          // [2Mp]
  void* cmod;           // bf16 [8][B][512] modulate vectors of the coordinate heads (shift | 1 + scale), fused coordinate head
  int* flags;           // [4] 0: any cond distance non-zero, 1: NaN seen
  // DMT_WO_EQ only (edge buffers above are then sized per DIRECTED edge)
  float* pab;           // [Mn,128] node2edge_lin halves applied per atom: [W[:, :256] hn | W[:, 256:] hn]
  void* eb;             // act [2Mp,64] copy of the edge stream
  float* pred_dir;      // [2Mp,2] edge-head output per directed edge (symmetrised afterwards)
};

size_t denoise_ws_carve(Arena& a, DenoiseWs& w, int B, int Mn, int Mp, bool bf16mode, int model_kind);

// ----------------------------------------------------------------------------- launchers
int linear(DsContext* ctx, const void* A, int lda, const void* W, int ldw, const float* bias, const float* addmat,
           int ldadd, void* out, int ldo, int out_dtype, int M, int N, int K, int act, cudaStream_t s);

// per-molecule noise level: either nl[B] (forward API) or coef[(*step)*4+3] broadcast (sampling loop)
struct StepRef {
  const float* coef;    // [steps,4] (c_x, c_pred, sigma, noise_level) or null
  const int* step;      // device step counter or null
};

int denoise_packed(DsContext* ctx, const PackedWeights& pw, const Plan& plan, const float* xs, const float* es,
                   const float* cond_x, const float* cond_e, const float* noise_level, StepRef sr,
                   const float* ctx_emb, float* pred_x, float* pred_e, DenoiseWs& w, cudaStream_t s);
// DMT_WO_EQ (dmt_wo_eq_kernels.cu); same contract
int denoise_wo_eq_packed(DsContext* ctx, const PackedWeights& pw, const Plan& plan, const float* xs, const float* es,
                         const float* cond_x, const float* cond_e, const float* noise_level, StepRef sr,
                         const float* ctx_emb, float* pred_x, float* pred_e, DenoiseWs& w, cudaStream_t s);

// fused coordinate head of one block (coord_head_tc.cu): pair part of input_lin -> LayerNorm + modulate -> coord_mlp ->
// w[d], on CTA pairs (cta_group::2 MMAs); X = the [dist | e] operand [Mp,128] bf16, ab = hoisted per-atom parts [Mn,512] bf16
// cmod_l = this block's [B][512] bf16 modulate vectors (shift | 1 + scale), written once per call by coord_mod_launch
int coord_head_launch(DsContext* ctx, const Plan& plan, const void* X, const void* ab, const void* cmod_l, const uint8_t* pflags,
                      const void* we, const void* wc1, const float* bc1, const float* wc2, float* wdir, cudaStream_t s);
int coord_mod_launch(DsContext* ctx, int B, int n_blocks, const float* ada, void* cmod, cudaStream_t s);   // cmod [n_blocks][B][512] bf16
// one cta_group::2 MMA tile: out[256,256] f32 = A[256,K] W[256,K]^T (test probe of the CTA-pair operand split)
int umma2_probe_launch(DsContext* ctx, const void* A, const void* W, float* out, int K, cudaStream_t s);

// fused edge-stream update of one block (edge_ffn_tc.cu): residual + LN + modulate -> ff3 -> SiLU -> ff4 -> gated residual
// ws / bs / skip / lds: optional skip projection edge_l(e) [64 -> 16] of the updated rows into the edge-head operand (null: none)
int edge_ffn_launch(DsContext* ctx, const Plan& plan, float* e, void* xe, int ldx, const float* pn, const float* n2e_b,
                    const float* ada_l, const void* w3, const float* b3, const void* w4, const float* b4, const void* ws,
                    const float* bs, void* skip, int lds, cudaStream_t s);

int launch_pack_dense(DsContext* ctx, const Plan& plan, const float* x, const float* ex, float* xs, float* es,
                      cudaStream_t s);
int launch_unpack_dense(DsContext* ctx, const Plan& plan, const float* xs, const float* es, float* x, float* ex,
                        cudaStream_t s);

struct NoiseSrc {
  // external (validation) noise: raw randn draws in the reference's shapes, step-major
  const float* raw_pos;   // [steps][B,N,3] or null
  const float* raw_h;     // [steps][B,N,6]
  const float* raw_e;     // [steps][B,2,N,N]
  unsigned long long seed;      // Philox key when raw_* are null
  long long gid_base;           // global id of molecule 0 of this shard (sharding-invariant noise)
  int philox_step_offset;       // added to the step index for the Philox counter (stand-alone sampler step)
  int raw_step_base;            // raw_* arrays start at this step (segmented loop)
};
int launch_sampler_step(DsContext* ctx, const Plan& plan, float* xs, float* es, const float* pred_x,
                        const float* pred_e, float* xmean, float* emean, StepRef sr, int step_host,
                        const NoiseSrc& ns, float temperature, cudaStream_t s);
int launch_init_noise(DsContext* ctx, const Plan& plan, float* xs, float* es, const NoiseSrc& ns, cudaStream_t s);
int launch_step_inc(DsContext* ctx, int* step, cudaStream_t s);
int launch_post_process(DsContext* ctx, const Plan& plan, const float* xs, const float* es, float* pos, int* atom_type,
                        int* fc, float* bond, cudaStream_t s);

int launch_molecule_records(DsContext* ctx, const Plan& plan, const float* x_mean, const float* edge_mean, uint8_t* rec,
                            int rec_n, int rec_bytes, cudaStream_t s);

struct SpecWs {
  float* z; void* zb; float* qkv; float* scores; void* att; float* o; void* f; float* head; void* hb;
};
size_t spec_ws_carve(Arena& a, SpecWs& w, int Bc, int q_len, bool bf16mode);
int specformer_ctx(DsContext* ctx, const PackedWeights& pw, const float* const* spectra, int B, float* ctx_out,
                   void* workspace, size_t ws_bytes, cudaStream_t s);

int pack_weights(DsContext* ctx, const char* const* names, const void* const* ptrs, int n, void* blob, size_t blob_bytes,
                 PackedWeights* out, cudaStream_t s);
size_t packed_weights_bytes(DsContext* ctx);
