// C-ABI entry points of the sampling hot path: weights, molecule plan, SpecFormer context, one denoiser call on
// dense (reference-shaped) tensors, the graph-captured sampling loop, a stand-alone sampler step and post_process.
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/diffspectra_b200.h"
#include "kernels.cuh"

namespace {

struct GraphKey {     // everything that is baked into the captured step graph
  const void *ws, *plan, *ctx_emb, *coef, *raw_pos, *raw_h, *raw_e, *pw_blob;
  int B, N, Mn, Mp;
  unsigned long long seed;
  long long gid_base;
  float temperature;
  int first_step;   // external-noise indexing is relative to the segment start
};
struct CtxFull : DsContext {
  PackedWeights pw;
  GraphKey gkey;
};
// DsContext is allocated as CtxFull in api_core.cu (ds_create) — see ds_ctx_alloc below.

inline CtxFull* full(ds_ctx* h) { return reinterpret_cast<CtxFull*>(h); }

// Last kernel of the body of the WHILE node that holds one sampling step: the loop goes on while the device-side step
// counter (incremented by the step itself) is below the end of the segment.  step[0] = current step, step[1] = end.
__global__ void k_loop_condition(cudaGraphConditionalHandle handle, const int* __restrict__ step) {
  if (threadIdx.x == 0) cudaGraphSetConditional(handle, step[0] < step[1] ? 1u : 0u);
}

// plan blob layout (device, int32): n_atoms[B] | noff[B+1] | poff[B+1] | node_info[Mn_max] | pair_info[Mp_max] |
// dir_info[2*Mp_max] (int4) | dir_mol[2*Mp_max] | pair_rows[Mp_max] (int2) | mol_order[B] | node_order[Mn_max] | mol_launch[B] (int4) | atom_launch[Mn_max] (int4)
struct PlanLayout {
  size_t n_atoms, noff, poff, node_info, pair_info, dir_info, dir_mol, pair_rows, mol_order, node_order, mol_launch, atom_launch, total;
};
PlanLayout plan_layout(int B, int N) {
  PlanLayout L;
  auto al = [](size_t x) { return (x + 63) & ~size_t(63); };
  size_t o = 0;
  L.n_atoms = o; o = al(o + size_t(B) * 4);
  L.noff = o; o = al(o + size_t(B + 1) * 4);
  L.poff = o; o = al(o + size_t(B + 1) * 4);
  L.node_info = o; o = al(o + size_t(B) * N * 4);
  L.pair_info = o; o = al(o + size_t(B) * N * (N - 1) / 2 * 4 + 4);
  L.dir_info = o; o = al(o + size_t(B) * N * (N - 1) * 16 + 16);
  L.dir_mol = o; o = al(o + size_t(B) * N * (N - 1) * 4 + 4);
  L.pair_rows = o; o = al(o + size_t(B) * N * (N - 1) / 2 * 8 + 8);
  L.mol_order = o; o = al(o + size_t(B) * 4);
  L.node_order = o; o = al(o + size_t(B) * N * 4);
  L.mol_launch = o; o = al(o + size_t(B) * 16);
  L.atom_launch = o; o = al(o + size_t(B) * N * 16);
  L.total = o;
  return L;
}

int make_plan(const void* plan_dev, int B, int N, int Mn, int Mp, Plan* p) {
  DS_CHECK(plan_dev != nullptr, DS_ERR_INVALID, "plan buffer is null");
  const PlanLayout L = plan_layout(B, N);
  const uint8_t* base = static_cast<const uint8_t*>(plan_dev);
  p->B = B; p->N = N; p->Mn = Mn; p->Mp = Mp;
  p->n_atoms = reinterpret_cast<const int*>(base + L.n_atoms);
  p->noff = reinterpret_cast<const int*>(base + L.noff);
  p->poff = reinterpret_cast<const int*>(base + L.poff);
  p->node_info = reinterpret_cast<const uint32_t*>(base + L.node_info);
  p->pair_info = reinterpret_cast<const uint32_t*>(base + L.pair_info);
  p->dir_info = reinterpret_cast<const int4*>(base + L.dir_info);
  p->dir_mol = reinterpret_cast<const uint32_t*>(base + L.dir_mol);
  p->pair_rows = reinterpret_cast<const int2*>(base + L.pair_rows);
  p->mol_order = reinterpret_cast<const int*>(base + L.mol_order);
  p->node_order = reinterpret_cast<const int*>(base + L.node_order);
  p->mol_launch = reinterpret_cast<const int4*>(base + L.mol_launch);
  p->atom_launch = reinterpret_cast<const int4*>(base + L.atom_launch);
  return DS_OK;
}

struct LoopWs {
  float *xs, *es, *pred_x, *pred_e, *xmean, *emean;
  int* step;
};
size_t loop_ws_carve(Arena& a, LoopWs& w, int Mn, int Mp) {
  const size_t mn = Mn > 0 ? Mn : 1, mp = Mp > 0 ? Mp : 1;
  const size_t start = a.off;
  w.xs = static_cast<float*>(a.take(mn * 9 * 4));
  w.es = static_cast<float*>(a.take(mp * 2 * 4));
  w.pred_x = static_cast<float*>(a.take(mn * 9 * 4));
  w.pred_e = static_cast<float*>(a.take(mp * 2 * 4));
  w.xmean = static_cast<float*>(a.take(mn * 9 * 4));
  w.emean = static_cast<float*>(a.take(mp * 2 * 4));
  w.step = static_cast<int*>(a.take(16));
  return a.off - start;
}

}  // namespace

DsContext* ds_ctx_alloc() { return new CtxFull(); }
void ds_ctx_free(DsContext* c) { delete static_cast<CtxFull*>(c); }

extern "C" {

size_t ds_packed_weights_bytes(ds_ctx* h) { return packed_weights_bytes(full(h)); }

int ds_pack_weights(ds_ctx* h, const char* const* names, const void* const* ptrs, int n, void* blob, size_t blob_bytes,
                    void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c != nullptr && names != nullptr && ptrs != nullptr, DS_ERR_INVALID, "ds_pack_weights: null argument");
  c->pw.valid = false;
  if (c->step_graph) {   // packed pointers may have moved: drop the cached step graph
    cudaGraphExecDestroy(c->step_graph);
    c->step_graph = nullptr;
  }
  return pack_weights(c, names, ptrs, n, blob, blob_bytes, &c->pw, reinterpret_cast<cudaStream_t>(stream));
}

size_t ds_plan_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return plan_layout(B, N).total;
}

// Host side of the plan: every table of the blob (layout above) from the atom counts.  No CUDA call in here, so the
// CPU tests can check it (ds_plan_build_host) against an independent restatement.
static int fill_plan_host(const int* n_atoms_host, int B, int N, std::vector<uint8_t>& host, int* Mn_out, int* Mp_out) {
  DS_CHECK(n_atoms_host && Mn_out && Mp_out, DS_ERR_INVALID, "ds_plan_build: null argument");
  DS_CHECK(B > 0 && B < (1 << 19), DS_ERR_INVALID, "ds_plan_build: B=%d out of range", B);
  DS_CHECK(N > 0 && N <= MAX_ATOMS, DS_ERR_INVALID, "ds_plan_build: N=%d out of range (1..%d)", N, MAX_ATOMS);
  const PlanLayout L = plan_layout(B, N);
  host.assign(L.total, 0);
  int* na = reinterpret_cast<int*>(host.data() + L.n_atoms);
  int* noff = reinterpret_cast<int*>(host.data() + L.noff);
  int* poff = reinterpret_cast<int*>(host.data() + L.poff);
  uint32_t* ni = reinterpret_cast<uint32_t*>(host.data() + L.node_info);
  uint32_t* pi = reinterpret_cast<uint32_t*>(host.data() + L.pair_info);
  int4* di = reinterpret_cast<int4*>(host.data() + L.dir_info);
  uint32_t* dm = reinterpret_cast<uint32_t*>(host.data() + L.dir_mol);
  int2* prw = reinterpret_cast<int2*>(host.data() + L.pair_rows);
  int mn = 0, mp = 0;
  for (int b = 0; b < B; ++b) {
    const int n = n_atoms_host[b];
    DS_CHECK(n >= 1 && n <= N, DS_ERR_INVALID, "ds_plan_build: n_atoms[%d]=%d outside 1..%d", b, n, N);
    na[b] = n;
    noff[b] = mn;
    poff[b] = mp;
    for (int i = 0; i < n; ++i) ni[mn + i] = (static_cast<uint32_t>(b) << 6) | i;
    int q = mp;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) {
        prw[q] = make_int2(mn + i, mn + j);
        pi[q++] = (static_cast<uint32_t>(b) << 12) | (i << 6) | j;
      }
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) {
        if (c == r) continue;
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        const int pr = mp + lo * n - (lo * (lo + 1)) / 2 + (hi - lo - 1);
        const size_t dd = static_cast<size_t>(2) * mp + static_cast<size_t>(r) * (n - 1) + (c - (c > r ? 1 : 0));
        di[dd] = make_int4(pr, mn + r, mn + c, b);
        dm[dd] = static_cast<uint32_t>(b);
      }
    mn += n;
    mp += n * (n - 1) / 2;
  }
  noff[B] = mn;
  poff[B] = mp;
  {
    // largest molecules first: the per-molecule kernels (cost ~ n^2) then end on their cheapest CTAs
    int* ord = reinterpret_cast<int*>(host.data() + L.mol_order);
    std::vector<int> idx(B);
    for (int b = 0; b < B; ++b) idx[b] = b;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return na[a] > na[b]; });
    for (int b = 0; b < B; ++b) ord[b] = idx[b];
    // the same order atom by atom, for the kernels whose warps map to atoms
    int* nord = reinterpret_cast<int*>(host.data() + L.node_order);
    int4* ml = reinterpret_cast<int4*>(host.data() + L.mol_launch);
    int4* al4 = reinterpret_cast<int4*>(host.data() + L.atom_launch);
    int q = 0;
    for (int b = 0; b < B; ++b) {
      const int m = idx[b];
      ml[b] = make_int4(m, na[m], noff[m], poff[m]);
      for (int i = 0; i < na[m]; ++i) {
        al4[q] = make_int4(noff[m] + i, m, (na[m] << 8) | i, poff[m]);
        nord[q++] = noff[m] + i;
      }
    }
  }
  *Mn_out = mn;
  *Mp_out = mp;
  return DS_OK;
}


int ds_plan_build(ds_ctx* h, const int* n_atoms_host, int B, int N, void* plan_dev, int* Mn_out, int* Mp_out, void* stream) {
  DS_CHECK(h && plan_dev, DS_ERR_INVALID, "ds_plan_build: null argument");
  std::vector<uint8_t> host;
  DS_TRY(fill_plan_host(n_atoms_host, B, N, host, Mn_out, Mp_out));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DS_CUDA_CHECK(cudaMemcpyAsync(plan_dev, host.data(), host.size(), cudaMemcpyHostToDevice, s));
  DS_CUDA_CHECK(cudaStreamSynchronize(s));   // host staging buffer dies at return
  return DS_OK;
}

int ds_plan_build_host(const int* n_atoms_host, int B, int N, void* plan_host, size_t plan_bytes, int* Mn_out, int* Mp_out) {
  DS_CHECK(plan_host != nullptr, DS_ERR_INVALID, "ds_plan_build_host: null argument");
  std::vector<uint8_t> host;
  DS_TRY(fill_plan_host(n_atoms_host, B, N, host, Mn_out, Mp_out));
  DS_CHECK(plan_bytes >= host.size(), DS_ERR_INVALID, "ds_plan_build_host: buffer of %zu bytes, need %zu", plan_bytes, host.size());
  memcpy(plan_host, host.data(), host.size());
  return DS_OK;
}

int ds_plan_layout(int B, int N, size_t* offsets, int n_offsets) {
  // byte offsets of the tables inside the plan blob, in declaration order (tests / debugging)
  DS_CHECK(offsets != nullptr && n_offsets >= 13 && B > 0 && N > 0, DS_ERR_INVALID, "ds_plan_layout: need room for 13 offsets");
  const PlanLayout L = plan_layout(B, N);
  const size_t o[13] = {L.n_atoms, L.noff, L.poff, L.node_info, L.pair_info, L.dir_info, L.dir_mol, L.pair_rows, L.mol_order,
                        L.node_order, L.mol_launch, L.atom_launch, L.total};
  for (int i = 0; i < 13; ++i) offsets[i] = o[i];
  return DS_OK;
}

size_t ds_workspace_bytes(ds_ctx* h, int B, int Mn, int Mp) {
  CtxFull* c = full(h);
  if (!c || B <= 0) return 0;
  Arena a{nullptr, 0, 0, true};
  DenoiseWs dw;
  LoopWs lw;
  denoise_ws_carve(a, dw, B, Mn, Mp, ds_is_bf16(c), c->model_kind);
  loop_ws_carve(a, lw, Mn, Mp);
  a.take(size_t(B) * D_TIME * 4);   // ctx embedding when the loop computes it
  return a.off + 1024;
}

size_t ds_specformer_workspace_bytes(ds_ctx* h, int B) {
  CtxFull* c = full(h);
  if (!c || B <= 0) return 0;
  static const int kLen[3] = {701, 3501, 3501}, kPatch[3] = {20, 50, 50}, kStride[3] = {10, 25, 25};
  int Q = 0;
  for (int t = 0; t < 3; ++t)
    if (c->spectra_version == 3 || c->spectra_version == t) Q += (kLen[t] - kPatch[t]) / kStride[t] + 1;
  int Bc = static_cast<int>((size_t(1) << 30) / (size_t(16) * Q * Q * 4));
  if (Bc < 1) Bc = 1;
  if (Bc > B) Bc = B;
  Arena a{nullptr, 0, 0, true};
  SpecWs w;
  spec_ws_carve(a, w, Bc, Q, ds_is_bf16(c));
  return a.off + 1024;
}

int ds_specformer_ctx(ds_ctx* h, const float* uv, const float* ir, const float* raman, int B, float* ctx_out,
                      void* workspace, size_t workspace_bytes, void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c && ctx_out && workspace, DS_ERR_INVALID, "ds_specformer_ctx: null argument");
  const float* all[3] = {uv, ir, raman};
  const float* used[3] = {nullptr, nullptr, nullptr};
  if (c->spectra_version == 3) {
    used[0] = uv; used[1] = ir; used[2] = raman;
  } else {
    used[0] = all[c->spectra_version];
  }
  return specformer_ctx(c, c->pw, used, B, ctx_out, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int ds_denoise(ds_ctx* h, const void* plan_dev, int B, int N, int Mn, int Mp, const float* x, const float* edge_x,
               const float* cond_x, const float* cond_edge_x, const float* noise_level, const float* ctx_emb,
               float* out_x, float* out_edge, void* workspace, size_t workspace_bytes, void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c && x && edge_x && noise_level && ctx_emb && out_x && out_edge && workspace, DS_ERR_INVALID,
           "ds_denoise: null argument");
  Plan plan;
  DS_TRY(make_plan(plan_dev, B, N, Mn, Mp, &plan));
  Arena a{static_cast<uint8_t*>(workspace), 0, workspace_bytes, false};
  DenoiseWs dw;
  LoopWs lw;
  denoise_ws_carve(a, dw, B, Mn, Mp, ds_is_bf16(c), c->model_kind);
  loop_ws_carve(a, lw, Mn, Mp);
  DS_CHECK(a.off <= workspace_bytes, DS_ERR_WORKSPACE, "ds_denoise: workspace too small (%zu < %zu)", workspace_bytes, a.off);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DS_TRY(launch_pack_dense(c, plan, x, edge_x, lw.xs, lw.es, s));
  const float* cx = nullptr;
  const float* ce = nullptr;
  if (cond_x) {
    DS_TRY(launch_pack_dense(c, plan, cond_x, cond_edge_x, lw.xmean, lw.emean, s));
    cx = lw.xmean;
    ce = lw.emean;
  }
  StepRef sr{nullptr, nullptr};
  DS_TRY(denoise_packed(c, c->pw, plan, lw.xs, lw.es, cx, ce, noise_level, sr, ctx_emb, lw.pred_x, lw.pred_e, dw, s));
  DS_TRY(launch_unpack_dense(c, plan, lw.pred_x, lw.pred_e, out_x, out_edge, s));
  return DS_OK;
}

int ds_sample_loop(ds_ctx* h, const void* plan_dev, int B, int N, int Mn, int Mp, const float* z, const float* edge_z,
                   const float* ctx_emb, const float* coef_table, int first_step, int steps, const float* raw_pos,
                   const float* raw_h, const float* raw_e, unsigned long long seed, long long gid_base, float temperature,
                   int use_graph, float* x_mean_out, float* edge_mean_out, void* workspace, size_t workspace_bytes,
                   void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c && ctx_emb && coef_table && x_mean_out && edge_mean_out && workspace, DS_ERR_INVALID,
           "ds_sample_loop: null argument");
  DS_CHECK(steps > 0 && first_step >= 0, DS_ERR_INVALID, "ds_sample_loop: steps must be positive, first_step >= 0");
  DS_CHECK((raw_pos == nullptr) == (raw_h == nullptr) && (raw_pos == nullptr) == (raw_e == nullptr), DS_ERR_INVALID,
           "ds_sample_loop: raw_pos/raw_h/raw_e must all be given (external noise) or all null (Philox)");
  DS_CHECK((z == nullptr) == (edge_z == nullptr), DS_ERR_INVALID, "ds_sample_loop: z and edge_z go together");
  Plan plan;
  DS_TRY(make_plan(plan_dev, B, N, Mn, Mp, &plan));
  Arena a{static_cast<uint8_t*>(workspace), 0, workspace_bytes, false};
  DenoiseWs dw;
  LoopWs lw;
  denoise_ws_carve(a, dw, B, Mn, Mp, ds_is_bf16(c), c->model_kind);
  loop_ws_carve(a, lw, Mn, Mp);
  DS_CHECK(a.off <= workspace_bytes, DS_ERR_WORKSPACE, "ds_sample_loop: workspace too small (%zu < %zu)", workspace_bytes, a.off);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // external noise arrays hold the draws of THIS segment only: raw index = step - first_step
  NoiseSrc ns{raw_pos, raw_h, raw_e, seed, gid_base, 0, first_step};

  if (first_step == 0) {
    // initial state: supplied z_T (reference: sampling.py:442-447) or Philox draw index -1
    if (z) {
      DS_TRY(launch_pack_dense(c, plan, z, edge_z, lw.xs, lw.es, s));
    } else {
      NoiseSrc ns0{nullptr, nullptr, nullptr, seed, gid_base, 0, 0};
      DS_TRY(launch_init_noise(c, plan, lw.xs, lw.es, ns0, s));
    }
    // self-conditioning carry starts at zero == the reference's cond_x=None branch (dmt.py:332-335)
    DS_CUDA_CHECK(cudaMemsetAsync(lw.pred_x, 0, size_t(Mn) * 9 * 4, s));
    DS_CUDA_CHECK(cudaMemsetAsync(lw.pred_e, 0, size_t(Mp > 0 ? Mp : 1) * 2 * 4, s));
  }   // else: continue from the state a previous segment left in the workspace
  // device-side loop state: step[0] = current step, step[1] = end of this segment.  Pageable host memory: the copy is
  // staged by the runtime before cudaMemcpyAsync returns, so the stack array may die at return.
  const int loop_state[2] = {first_step, first_step + steps};
  DS_CUDA_CHECK(cudaMemcpyAsync(lw.step, loop_state, sizeof(loop_state), cudaMemcpyHostToDevice, s));
  StepRef sr{coef_table, lw.step};

  auto one_step = [&]() -> int {
    DS_TRY(denoise_packed(c, c->pw, plan, lw.xs, lw.es, lw.pred_x, lw.pred_e, nullptr, sr, ctx_emb, lw.pred_x, lw.pred_e, dw, s));
    DS_TRY(launch_sampler_step(c, plan, lw.xs, lw.es, lw.pred_x, lw.pred_e, lw.xmean, lw.emean, sr, 0, ns, temperature, s));
    DS_TRY(launch_step_inc(c, lw.step, s));
    return DS_OK;
  };

  if (!use_graph) {
    for (int i = 0; i < steps; ++i) DS_TRY(one_step());
  } else {
    // The whole loop is ONE graph launch: a WHILE conditional node whose body is one captured step (all kernels read the
    // step index from device memory; the body's last kernel re-arms the condition while step < end).  DS_LOOP_GRAPH=0
    // falls back to replaying a one-step graph `steps` times from the host.
    GraphKey key;
    memset(&key, 0, sizeof(key));
    key.ws = workspace; key.plan = plan_dev; key.ctx_emb = ctx_emb; key.coef = coef_table;
    key.raw_pos = raw_pos; key.raw_h = raw_h; key.raw_e = raw_e; key.pw_blob = c->pw.w_ada;
    key.B = B; key.N = N; key.Mn = Mn; key.Mp = Mp; key.seed = seed; key.gid_base = gid_base; key.temperature = temperature;
    key.first_step = raw_pos ? first_step : 0;
    const bool whole_loop = c->loop_graph != 0;
    if (c->step_graph == nullptr || memcmp(&c->gkey, &key, sizeof(key)) != 0 || c->step_graph_is_loop != whole_loop) {
      if (c->step_graph) {
        cudaGraphExecDestroy(c->step_graph);
        c->step_graph = nullptr;
      }
      const long long before = c->launch_count;
      cudaGraph_t graph = nullptr;
      // The legacy default stream cannot be captured: record the step on a private stream, replay on the caller's.
      if (c->capture_stream == nullptr) DS_CUDA_CHECK(cudaStreamCreateWithFlags(&c->capture_stream, cudaStreamNonBlocking));
      cudaStream_t user_stream = s;
      s = c->capture_stream;
      int r = DS_OK;
      cudaError_t ce = cudaSuccess;
      if (whole_loop) {
        DS_CUDA_CHECK(cudaGraphCreate(&graph, 0));
        cudaGraphConditionalHandle handle;
        DS_CUDA_CHECK(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t node;
        DS_CUDA_CHECK(cudaGraphAddNode(&node, graph, nullptr, 0, &np));
        cudaGraph_t body = np.conditional.phGraph_out[0];
        DS_CUDA_CHECK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        r = one_step();
        if (r == DS_OK) {
          k_loop_condition<<<1, 32, 0, s>>>(handle, lw.step);
          c->launch_count++;
        }
        ce = cudaStreamEndCapture(s, nullptr);
      } else {
        DS_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        r = one_step();
        ce = cudaStreamEndCapture(s, &graph);
      }
      s = user_stream;
      if (r != DS_OK) {
        if (graph) cudaGraphDestroy(graph);
        return r;
      }
      if (ce != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        DS_CUDA_CHECK(ce);
      }
      c->step_graph_launches = c->launch_count - before;
      c->launch_count = before;
      cudaError_t ie = cudaGraphInstantiate(&c->step_graph, graph, 0);
      cudaGraphDestroy(graph);
      DS_CUDA_CHECK(ie);
      c->gkey = key;
      c->step_graph_is_loop = whole_loop;
    }
    if (whole_loop) {
      DS_CUDA_CHECK(cudaGraphLaunch(c->step_graph, s));
      c->launch_count += c->step_graph_launches * steps;
    } else {
      for (int i = 0; i < steps; ++i) {
        DS_CUDA_CHECK(cudaGraphLaunch(c->step_graph, s));
        c->launch_count += c->step_graph_launches;
      }
    }
  }
  // the reference returns the MEANS of the last step (sampling.py:628-629)
  DS_TRY(launch_unpack_dense(c, plan, lw.xmean, lw.emean, x_mean_out, edge_mean_out, s));
  return DS_OK;
}

int ds_sampler_step(ds_ctx* h, const void* plan_dev, int B, int N, int Mn, int Mp, float* x, float* edge_x,
                    const float* pred, const float* edge_pred, const float* coef_row, const float* raw_pos,
                    const float* raw_h, const float* raw_e, unsigned long long seed, long long gid_base, int step_index,
                    float temperature, float* x_mean_out, float* edge_mean_out, void* workspace, size_t workspace_bytes,
                    void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c && x && edge_x && pred && edge_pred && coef_row && workspace, DS_ERR_INVALID, "ds_sampler_step: null argument");
  Plan plan;
  DS_TRY(make_plan(plan_dev, B, N, Mn, Mp, &plan));
  Arena a{static_cast<uint8_t*>(workspace), 0, workspace_bytes, false};
  LoopWs lw;
  loop_ws_carve(a, lw, Mn, Mp);
  DS_CHECK(a.off <= workspace_bytes, DS_ERR_WORKSPACE, "ds_sampler_step: workspace too small (%zu < %zu)", workspace_bytes, a.off);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DS_TRY(launch_pack_dense(c, plan, x, edge_x, lw.xs, lw.es, s));
  DS_TRY(launch_pack_dense(c, plan, pred, edge_pred, lw.pred_x, lw.pred_e, s));
  // coef_row points at the row of this step; external noise pointers at this step's draws -> index 0 for both,
  // while Philox is keyed by the true step_index.
  NoiseSrc ns{raw_pos, raw_h, raw_e, seed, gid_base, step_index, 0};
  StepRef sr{coef_row, nullptr};
  DS_TRY(launch_sampler_step(c, plan, lw.xs, lw.es, lw.pred_x, lw.pred_e, lw.xmean, lw.emean, sr, 0, ns, temperature, s));
  DS_TRY(launch_unpack_dense(c, plan, lw.xs, lw.es, x, edge_x, s));
  if (x_mean_out) DS_TRY(launch_unpack_dense(c, plan, lw.xmean, lw.emean, x_mean_out, edge_mean_out, s));
  return DS_OK;
}

int ds_post_process(ds_ctx* h, const void* plan_dev, int B, int N, int Mn, int Mp, const float* x_mean, const float* edge_mean,
                    float* pos, int* atom_type, int* formal_charge, float* bond, void* workspace, size_t workspace_bytes,
                    void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c && x_mean && edge_mean && pos && atom_type && formal_charge && bond && workspace, DS_ERR_INVALID,
           "ds_post_process: null argument");
  Plan plan;
  DS_TRY(make_plan(plan_dev, B, N, Mn, Mp, &plan));
  Arena a{static_cast<uint8_t*>(workspace), 0, workspace_bytes, false};
  LoopWs lw;
  loop_ws_carve(a, lw, Mn, Mp);
  DS_CHECK(a.off <= workspace_bytes, DS_ERR_WORKSPACE, "ds_post_process: workspace too small");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DS_TRY(launch_pack_dense(c, plan, x_mean, edge_mean, lw.xs, lw.es, s));
  return launch_post_process(c, plan, lw.xs, lw.es, pos, atom_type, formal_charge, bond, s);
}

size_t ds_record_bytes(int N) { return N > 0 ? static_cast<size_t>(N) * 14 + static_cast<size_t>(N) * N + 1 : 0; }

int ds_molecule_records(ds_ctx* h, const void* plan_dev, int B, int N, int Mn, int Mp, const float* x_mean, const float* edge_mean,
                        int rec_n, void* records, size_t records_bytes, void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c && x_mean && edge_mean && records, DS_ERR_INVALID, "ds_molecule_records: null argument");
  DS_CHECK(rec_n >= N && rec_n <= 255, DS_ERR_INVALID, "ds_molecule_records: rec_n=%d must be in [N=%d, 255]", rec_n, N);
  const size_t rb = ds_record_bytes(rec_n);
  DS_CHECK(records_bytes >= rb * static_cast<size_t>(B), DS_ERR_WORKSPACE, "ds_molecule_records: %zu bytes given, %zu needed", records_bytes,
           rb * static_cast<size_t>(B));
  Plan plan;
  DS_TRY(make_plan(plan_dev, B, N, Mn, Mp, &plan));
  return launch_molecule_records(c, plan, x_mean, edge_mean, static_cast<uint8_t*>(records), rec_n, static_cast<int>(rb),
                                 reinterpret_cast<cudaStream_t>(stream));
}

int ds_umma2_probe(ds_ctx* h, const void* A, const void* W, float* out, int K, void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c && A && W && out, DS_ERR_INVALID, "ds_umma2_probe: null argument");
  return umma2_probe_launch(c, A, W, out, K, reinterpret_cast<cudaStream_t>(stream));
}

int ds_coord_head(ds_ctx* h, const void* plan_dev, int B, int N, int Mn, int Mp, const void* X, const void* ab, const float* ada_block,
                  const unsigned char* pflags, const void* we, const void* wc1_half, const float* bc1_half, const float* wc2, float* wdir,
                  void* scratch, void* stream) {
  CtxFull* c = full(h);
  DS_CHECK(c && X && ab && ada_block && pflags && we && wc1_half && bc1_half && wc2 && wdir && scratch, DS_ERR_INVALID,
           "ds_coord_head: null argument");
  Plan plan;
  DS_TRY(make_plan(plan_dev, B, N, Mn, Mp, &plan));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DS_TRY(coord_mod_launch(c, B, 1, ada_block, scratch, s));      // bf16 modulate vectors of this one block
  return coord_head_launch(c, plan, X, ab, scratch, pflags, we, wc1_half, bc1_half, wc2, wdir, s);
}

}  // extern "C"
