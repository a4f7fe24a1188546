// C-ABI entry points: context lifetime, error reporting and the GEMM test hook.
#include <stdarg.h>
#include <stdio.h>

#include <map>
#include <vector>
#include <stdlib.h>

#include "../../include/diffspectra_b200.h"
#include "context.cuh"

DsContext* ds_ctx_alloc();
void ds_ctx_free(DsContext* c);

static thread_local char g_err[1024] = "";

void ds_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

DsProf g_ds_prof;
bool g_ds_use_pdl = false;     // measured on B200: no gain with the implicit trigger (graph replay gaps are already ~0), -3.5% with an early trigger

extern "C" {

const char* ds_last_error(void) { return g_err; }

int ds_version(void) { return 100; }

int ds_create(ds_ctx** out, int device, int mode, int spectra_version) {
  return ds_create_model(out, device, mode, spectra_version, DS_MODEL_DMT);
}

int ds_create_model(ds_ctx** out, int device, int mode, int spectra_version, int model_kind) {
  DS_CHECK(out != nullptr, DS_ERR_INVALID, "ds_create: null out");
  DS_CHECK(model_kind == DS_MODEL_DMT || model_kind == DS_MODEL_DMT_WO_EQ, DS_ERR_INVALID, "ds_create: unknown model kind %d", model_kind);
  DS_CHECK(mode == 0 || mode == 1, DS_ERR_INVALID, "ds_create: mode must be 0 (fp32) or 1 (bf16), got %d", mode);
  DS_CHECK(spectra_version >= 0 && spectra_version <= 3, DS_ERR_INVALID, "ds_create: bad spectra_version %d",
           spectra_version);
  int ndev = 0;
  DS_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  DS_CHECK(device >= 0 && device < ndev, DS_ERR_INVALID, "ds_create: device %d out of range (%d visible)", device, ndev);
  DS_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  DS_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  DS_CHECK(prop.major == 10, DS_ERR_UNSUPPORTED,
           "ds_create: device %d is sm_%d%d; this library is built for sm_100a only (no fallback path)", device,
           prop.major, prop.minor);
  DsContext* c = ds_ctx_alloc();
  c->device = device;
  c->mode = mode;
  c->spectra_version = spectra_version;
  c->model_kind = model_kind;
  c->num_sms = prop.multiProcessorCount;
  if (const char* fm = getenv("DS_FUSE_MASK")) c->fuse_mask = atoi(fm);
  if (const char* pd = getenv("DS_PDL")) g_ds_use_pdl = atoi(pd) != 0;
  if (const char* ov = getenv("DS_OVERLAP")) c->overlap = atoi(ov);
  if (const char* fv = getenv("DS_FFN")) c->ffn_variant = atoi(fv);
  if (const char* ap = getenv("DS_ATT_P1")) c->att_p1 = atoi(ap);
  if (const char* tm = getenv("DS_TANH_MIX")) c->tanh_mix = atoi(tm);
  if (const char* lg = getenv("DS_LOOP_GRAPH")) c->loop_graph = atoi(lg);
  if (const char* ag = getenv("DS_ATT_G")) c->att_g = atoi(ag) == 4 ? 4 : 8;
  if (const char* sp = getenv("DS_SPLIT")) sscanf(sp, "%d,%d", &c->edge_cap, &c->node_cap);
  if (c->edge_cap < 1 || c->node_cap < 1 || c->edge_cap + c->node_cap > prop.multiProcessorCount) {
    c->edge_cap = prop.multiProcessorCount * 5 / 6;
    c->node_cap = prop.multiProcessorCount - c->edge_cap;
  }
  if (cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    ds_set_error("ds_create: cannot create the side stream / events");
    ds_ctx_free(c);
    return DS_ERR_CUDA;
  }
  int r = gemm_tc_init(c);
  if (r != DS_OK) {
    ds_ctx_free(c);
    return r;
  }
  *out = reinterpret_cast<ds_ctx*>(c);
  return DS_OK;
}

int ds_destroy(ds_ctx* h) {
  DsContext* c = reinterpret_cast<DsContext*>(h);
  if (!c) return DS_OK;
  if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
  if (c->capture_stream) cudaStreamDestroy(c->capture_stream);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  ds_ctx_free(c);
  return DS_OK;
}

int ds_profile_begin(void) {
  for (auto& r : g_ds_prof.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_ds_prof.recs.clear();
  g_ds_prof.on = true;
  return DS_OK;
}

int ds_profile_end(char* out, size_t out_bytes) {
  g_ds_prof.on = false;
  DS_CUDA_CHECK(cudaDeviceSynchronize());
  typedef std::pair<const void*, long long> Key;
  std::map<Key, std::pair<int, double>> agg;
  std::vector<Key> order;
  for (auto& r : g_ds_prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) ms = 0.f;
    const Key k(r.fn, r.tag);
    auto it = agg.find(k);
    if (it == agg.end()) { agg[k] = std::make_pair(1, static_cast<double>(ms)); order.push_back(k); }
    else { it->second.first++; it->second.second += ms; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_ds_prof.recs.clear();
  size_t off = 0;
  if (out && out_bytes) out[0] = 0;
  for (const Key& k : order) {
    const char* name = nullptr;
    if (cudaFuncGetName(&name, k.first) != cudaSuccess || !name) name = "?";
    const int n = snprintf(out ? out + off : nullptr, out && off < out_bytes ? out_bytes - off : 0, "%s\t%lld\t%d\t%.3f\n", name,
                           k.second, agg[k].first, agg[k].second * 1000.0);
    if (n < 0 || !out || off + n >= out_bytes) break;
    off += n;
  }
  return DS_OK;
}

long long ds_launch_count(ds_ctx* h) { return reinterpret_cast<DsContext*>(h)->launch_count; }

int ds_gemm(ds_ctx* h, int use_tensor_cores, const void* A, int lda, const void* W, int ldw, const float* bias,
            const float* addmat, int ldadd, void* out, int ldo, int M, int N, int K, int in_dtype, int out_dtype,
            int act, void* stream) {
  DsContext* c = reinterpret_cast<DsContext*>(h);
  DS_CHECK(c != nullptr, DS_ERR_INVALID, "ds_gemm: null ctx");
  GemmDesc g;
  g.A = A; g.W = W; g.bias = bias; g.addmat = addmat; g.out = out;
  g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = ldw; g.ldo = ldo; g.ldadd = ldadd;
  g.a_dtype = in_dtype; g.out_dtype = out_dtype; g.act = act;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (use_tensor_cores) return gemm_tc_launch(c, g, s);
  c->launch_count++;
  return gemm_simt_launch(g, c->mode == 1, s);
}

int ds_gemm_fused(ds_ctx* h, int mode, const void* A, int lda, const void* W, int ldw, const float* bias, int M, int N, int K,
                  const unsigned* row_info, int info_shift, const float* ada, int off_a, int off_b, const float* resid,
                  int ldres, void* out, int ldo, void* out2, int ldo2, const float* wc2, const unsigned char* dflags,
                  float* wdir, void* stream) {
  DsContext* c = reinterpret_cast<DsContext*>(h);
  DS_CHECK(c != nullptr, DS_ERR_INVALID, "ds_gemm_fused: null ctx");
  GemmDesc g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.M = M; g.N = N; g.K = K;
  g.a_dtype = DT_BF16; g.mode = mode;
  g.out_dtype = (mode == GEMM_LNMOD) ? DT_BF16 : DT_F32;
  g.row_info = row_info; g.info_shift = info_shift; g.ada = ada; g.off_a = off_a; g.off_b = off_b;
  g.resid = resid; g.ldres = ldres; g.out = out; g.ldo = ldo; g.out2 = out2; g.ldo2 = ldo2;
  g.wc2 = wc2; g.pflags = dflags; g.wdir = wdir;
  return gemm_tc_launch(c, g, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
