// Host-side context object behind the opaque ds_ctx handle of the C-ABI.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

struct PackedWeights;   // weights.cuh

struct DsContext {
  int device = 0;
  int mode = 0;              // 0 = fp32 validation (SIMT GEMM, libm maths), 1 = bf16 production (tcgen05 GEMM)
  int spectra_version = 3;   // 0 uv, 1 ir, 2 raman, 3 allspectra
  int num_sms = 148;
  void* encode_tiled = nullptr;      // cuTensorMapEncodeTiled (driver entry point, resolved at run time)
  int fuse_mask = 15;                // debug: bit0 LNMOD, bit1 RESGATE(node), bit2 RESGATE(edge), bit3 COORD epilogue, bit4 fused coordinate head (coord_tc.cu; off: measured slower than k_coord_ln + COORD GEMM)
  long long launch_count = 0;        // kernels launched (or captured) through this context
  // cached CUDA graph of one sampling step (ds_sample_loop)
  cudaGraphExec_t step_graph = nullptr;
  cudaStream_t capture_stream = nullptr;
  long long step_graph_launches = 0;   // kernels inside one replay of step_graph
};

inline bool ds_is_bf16(const DsContext* c) { return c->mode == 1; }
