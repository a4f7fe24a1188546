/* diffspectra_b200 — C-ABI of the B200-native (sm_100a) DiffSpectra reverse-diffusion sampling hot path.
 *
 * This is the drop-in boundary: every entry point below replaces a piece of the reference's
 * PyTorch-eager path and is what a binding in the reference's host language (Python -> ctypes, see
 * INTEGRATION.md and diffspectra_b200/_lib.py) calls.  Conventions:
 *   - plain pointers and sizes only; every data pointer is DEVICE memory unless the name ends in `_host`;
 *   - the caller owns every buffer (sizes from the ds_*_bytes queries); no entry point allocates device memory;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises the device
 *     except ds_plan_build (host staging buffer);
 *   - return 0 on success, a negative DS_ERR_* code otherwise (never throws); ds_last_error() gives the message;
 *   - there is NO CPU fallback: on a non-sm_100 device ds_create fails with DS_ERR_UNSUPPORTED.
 *
 * Shapes follow the reference: B molecules, N padded atoms, x = [pos(3) | atom type(5) | formal charge(1)],
 * edge_x = [bond exists | bond order] (configs/diffspectra_qm9s.py:26-30,51-57).  All floating tensors are fp32,
 * contiguous.  Edge tensors must be symmetric in (i, j) with a zero diagonal, as they always are on the sampling
 * path (models/utils.py:100-106, models/dmt.py:399).
 */
#ifndef DIFFSPECTRA_B200_H
#define DIFFSPECTRA_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_OK 0
#define DS_ERR_INVALID (-1)       /* bad argument (reference: Python ValueError / AssertionError) */
#define DS_ERR_CUDA (-2)          /* CUDA runtime / driver error */
#define DS_ERR_MISSING_PARAM (-3) /* state_dict lacks a parameter (reference: load_state_dict strict=True KeyError) */
#define DS_ERR_UNSUPPORTED (-4)   /* device is not sm_100 */
#define DS_ERR_WORKSPACE (-5)     /* caller buffer too small */

#define DS_MODE_FP32 0 /* validation: CUDA-core fp32 GEMMs, libm transcendentals (1e-5 parity gate) */
#define DS_MODE_BF16 1 /* production: tcgen05 bf16 GEMMs (fp32 accumulate), SFU transcendentals */

#define DS_SPECTRA_UV 0
#define DS_SPECTRA_IR 1
#define DS_SPECTRA_RAMAN 2
#define DS_SPECTRA_ALL 3

#define DS_MODEL_DMT 0       /* models/dmt.py DMT (equivariant) */
#define DS_MODEL_DMT_WO_EQ 1 /* models/dmt_wo_eq.py DMT_WO_EQ (non-equivariant ablation) */

typedef struct ds_ctx ds_ctx;

const char* ds_last_error(void);
int ds_version(void);

/* Context = one model instance on one device.  Replaces models/utils.py:24-28 (create_model -> .to(device));
 * one context per process/GPU instead of nn.DataParallel replication. */
int ds_create(ds_ctx** out, int device, int mode, int spectra_version);
/* Same with the model family selected explicitly (registry names 'DMT' / 'DMT_WO_EQ', models/utils.py:5-21). */
int ds_create_model(ds_ctx** out, int device, int mode, int spectra_version, int model_kind);
int ds_destroy(ds_ctx* ctx);
/* kernels launched (or replayed from the captured step graph) through this context so far */
long long ds_launch_count(ds_ctx* ctx);

/* In-stream timing of every kernel launched through the library between the two calls (use ds_sample_loop with
 * use_graph = 0, or ds_denoise): ds_profile_end synchronises the device and writes one line per kernel,
 * "<mangled name>\t<tag>\t<launches>\t<total microseconds>\n", into out (tag = mode<<61 | N<<46 | K<<30 | M for the
 * tcgen05 GEMM, 0 otherwise).  Measurement aid for bench.py; not thread-safe. */
int ds_profile_begin(void);
int ds_profile_end(char* out, size_t out_bytes);

/* Weights.  `names[i]` / `ptrs[i]` = state_dict entries (fp32, contiguous, device) of a reference-compatible DMT
 * module, parameters AND BatchNorm buffers, names as in models/dmt.py:211-262 (an optional "module." prefix from
 * nn.DataParallel checkpoints, utils.py:15-19, is stripped).  Re-packs them into `blob` (GEMM-ready bf16/fp32,
 * adaLN projections stacked, q/k/v and edge0/edge1 fused).  Call again after the parameters change
 * (restore_checkpoint / ema.copy_to, run_lib.py:361-362). */
size_t ds_packed_weights_bytes(ds_ctx* ctx);
int ds_pack_weights(ds_ctx* ctx, const char* const* names, const void* const* ptrs, int n, void* blob,
                    size_t blob_bytes, void* stream);

/* Molecule plan: the packed ragged layout (atoms, unordered pairs) for a batch with n_atoms_host[B] valid atoms
 * each, padded to N in the dense interface.  Replaces the per-call adj_mask.nonzero() + dense_to_sparse()
 * (models/dmt.py:327-329) and the node/edge mask tensors (sampling.py:429-439). */
size_t ds_plan_bytes(int B, int N);
int ds_plan_build(ds_ctx* ctx, const int* n_atoms_host, int B, int N, void* plan, int* Mn_out, int* Mp_out,
                  void* stream);
/* The same blob into HOST memory (no CUDA call; what ds_plan_build copies to the device), and the byte offsets of its 12
 * tables + the total size, in order: n_atoms, noff, poff, node_info, pair_info, dir_info, dir_mol, pair_rows, mol_order,
 * node_order, mol_launch, atom_launch, total.  Used by the CPU tests of the host logic. */
int ds_plan_build_host(const int* n_atoms_host, int B, int N, void* plan_host, size_t plan_bytes, int* Mn_out, int* Mp_out);
int ds_plan_layout(int B, int N, size_t* offsets, int n_offsets);

size_t ds_workspace_bytes(ds_ctx* ctx, int B, int Mn, int Mp);
size_t ds_specformer_workspace_bytes(ds_ctx* ctx, int B);

/* ctx_out[B,1024] = cond_lin(SpecFormer(spectra))  — models/dmt.py:348-350 + models/specformer.py:77-120.
 * uv[B,701], ir[B,3501], raman[B,3501]; only the spectra of the context's spectra_version are read. */
int ds_specformer_ctx(ds_ctx* ctx, const float* uv, const float* ir, const float* raman, int B, float* ctx_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* One denoiser call on reference-shaped tensors — DMT.forward, models/dmt.py:306-413, called from
 * sampling.py:588-589.  x[B,N,9], edge_x[B,N,N,2], cond_x / cond_edge_x same shapes or both NULL (first step),
 * noise_level[B], ctx_emb[B,1024] from ds_specformer_ctx -> out_x[B,N,9], out_edge[B,N,N,2]
 * (padded rows / edges / diagonal exactly 0).  Inputs are not modified. */
int ds_denoise(ds_ctx* ctx, const void* plan, int B, int N, int Mn, int Mp, const float* x, const float* edge_x,
               const float* cond_x, const float* cond_edge_x, const float* noise_level, const float* ctx_emb,
               float* out_x, float* out_edge, void* workspace, size_t workspace_bytes, void* stream);

/* The whole reverse-diffusion loop — AncestralSampler.sampling, sampling.py:565-631 (pred_data, pred_edge,
 * self_cond='ori').  coef_table[steps,4] = (c_x, c_pred, sigma, noise_level) rows built by the host from
 * NoiseScheduleVP.marginal_prob with the reference's op order (sampling.py:571-584).
 * Runs table rows [first_step, first_step + steps).  first_step == 0 starts a new trajectory from z[B,N,9] /
 * edge_z[B,N,N,2] (or, if both NULL, from noise drawn on the device); first_step > 0 continues from the state the
 * previous segment left in `workspace` (same plan / workspace; z, edge_z ignored).
 * Noise per step: raw_pos[steps,B,N,3], raw_h[steps,B,N,6], raw_e[steps,B,2,N,N] = the reference's randn draws for
 * THIS segment (validation), or all NULL for device Philox keyed by (seed, gid_base + molecule, step) — invariant
 * to sharding and to segmentation.
 * use_graph != 0 captures one step as a CUDA graph and replays it `steps` times.
 * Outputs the MEANS of the last step (sampling.py:628-629): x_mean_out[B,N,9], edge_mean_out[B,N,N,2]. */
int ds_sample_loop(ds_ctx* ctx, const void* plan, int B, int N, int Mn, int Mp, const float* z, const float* edge_z,
                   const float* ctx_emb, const float* coef_table, int first_step, int steps, const float* raw_pos,
                   const float* raw_h, const float* raw_e, unsigned long long seed, long long gid_base, float temperature,
                   int use_graph, float* x_mean_out, float* edge_mean_out, void* workspace, size_t workspace_bytes,
                   void* stream);

/* One fused ancestral update (sampling.py:605-624) on reference-shaped tensors, in place on x / edge_x:
 * coef_row[4] = this step's (c_x, c_pred, sigma, .); raw_* = this step's randn draws ([B,N,3],[B,N,6],[B,2,N,N])
 * or NULL for Philox at step_index.  Optionally also returns the means. */
int ds_sampler_step(ds_ctx* ctx, const void* plan, int B, int N, int Mn, int Mp, float* x, float* edge_x,
                    const float* pred, const float* edge_pred, const float* coef_row, const float* raw_pos,
                    const float* raw_h, const float* raw_e, unsigned long long seed, long long gid_base, int step_index,
                    float temperature, float* x_mean_out, float* edge_mean_out, void* workspace, size_t workspace_bytes,
                    void* stream);

/* post_process — sampling.py:53-97 with the inverse scaler of utils.py:71-105 (factors 1,4,4,1; centered;
 * compress_edge): pos[B,N,3], atom_type[B,N] (argmax), formal_charge[B,N] (round(4 fc)), bond[B,N,N] (0..3). */
int ds_post_process(ds_ctx* ctx, const void* plan, int B, int N, int Mn, int Mp, const float* x_mean,
                    const float* edge_mean, float* pos, int* atom_type, int* formal_charge, float* bond,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Molecule records — mol_process + post_process (sampling.py:12-32,53-97) fused: from the reference-shaped means of the
 * last step straight to one fixed-size byte record per molecule,
 *   pos f32[R*3] | atom_type u8[R] | formal_charge i8[R] | bond u8[R*R] | n_atoms u8      (ds_record_bytes(R) = 14 R + R^2 + 1)
 * R = rec_n >= N (e.g. the data set's max_node, so that rounds with a smaller padded N emit records of one size), with
 * the same discretisation as ds_post_process; padded atoms / bonds are 0.  records[B * ds_record_bytes(rec_n)] is the unit
 * of the single D2H copy of the eval driver and of the one all-gather of the multi-GPU path. */
size_t ds_record_bytes(int rec_n);
int ds_molecule_records(ds_ctx* ctx, const void* plan, int B, int N, int Mn, int Mp, const float* x_mean,
                        const float* edge_mean, int rec_n, void* records, size_t records_bytes, void* stream);

/* Test hook: out[M,N] = act(A[M,K] W[N,K]^T + bias + addmat).  use_tensor_cores=1 -> the tcgen05/TMA kernel
 * (bf16 A/W), 0 -> the CUDA-core kernel.  dtype 0 = f32, 1 = bf16; act 0 none, 1 SiLU, 2 tanh, 3 GELU(erf). */
int ds_gemm(ds_ctx* ctx, int use_tensor_cores, const void* A, int lda, const void* W, int ldw, const float* bias,
            const float* addmat, int ldadd, void* out, int ldo, int M, int N, int K, int in_dtype, int out_dtype,
            int act, void* stream);

/* Test hook for the fused epilogues of the tcgen05 GEMM (bf16 A/W).  mode 1 = LNMOD: out(bf16)[M,64] =
 * modulate(LayerNorm(A W^T + bias), ada[mol][off_a:], ada[mol][off_b:]);  mode 2 = RESGATE: out(f32) = resid +
 * ada[mol][off_a:] * (A W^T + bias), out2(bf16) = copy;  mode 3 = COORD: wdir[row] = mean(tanh(wc2 . SiLU(2 (A W^T +
 * bias))) * [1, dflags bit0, dflags bit1]) (W, bias = the halved first layer).  mol = row_info[row] >> info_shift; ada rows have stride 19584 floats. */
int ds_gemm_fused(ds_ctx* ctx, int mode, const void* A, int lda, const void* W, int ldw, const float* bias, int M, int N,
                  int K, const unsigned* row_info, int info_shift, const float* ada, int off_a, int off_b,
                  const float* resid, int ldres, void* out, int ldo, void* out2, int ldo2, const float* wc2,
                  const unsigned char* dflags, float* wdir, void* stream);

/* Test hook for the CTA-pair (cta_group::2) tensor-core path: out[256,256] (f32) = A[256,K] W[256,K]^T (bf16, K in
 * {64,128,192,256}) computed by ONE cluster of two CTAs, each holding 128 rows of A and of W. */
int ds_umma2_probe(ds_ctx* ctx, const void* A, const void* W, float* out, int K, void* stream);

/* Test hook for the fused coordinate head of one block (models/dmt.py:37-60, csrc/coord_head_tc.cu): for every directed edge
 * d = (r -> c) of the plan, wdir[d] = mean(tanh(wc2 . SiLU(2 (wc1_half . z + bc1_half))) * [1, adj2d, adjsp]) with
 * z = modulate(LayerNorm(ab[r][0:256] + ab[c][256:512] + we . X[pair]), ada[mol][1920:2176], ada[mol][2176:2432]).
 * X[Mp,128], ab[Mn,512], we[256,128], wc1_half[256,256] bf16; ada_block = adaLN table of the block (row stride 19584 floats);
 * pflags[Mp] adjacency bits; wdir[2 Mp] source-major (d = 2 poff[mol] + r (n-1) + c - (c > r)); scratch = B * 1024 bytes
 * (bf16 copy of the modulate vectors). */
int ds_coord_head(ds_ctx* ctx, const void* plan, int B, int N, int Mn, int Mp, const void* X, const void* ab,
                  const float* ada_block, const unsigned char* pflags, const void* we, const void* wc1_half,
                  const float* bc1_half, const float* wc2, float* wdir, void* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFSPECTRA_B200_H */
