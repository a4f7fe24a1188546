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
  int model_kind = 0;        // 0 = DMT (models/dmt.py), 1 = DMT_WO_EQ ablation (models/dmt_wo_eq.py)
  int num_sms = 148;
  void* encode_tiled = nullptr;      // cuTensorMapEncodeTiled (driver entry point, resolved at run time)
  int fuse_mask = 511;                // debug: bit0 LNMOD, bit1 RESGATE(node), bit2 RESGATE(edge), bit3 COORD epilogue, bit4 fused coordinate head on CTA pairs (coord_head_tc.cu: replaces the gp GEMM, k_coord_ln and the COORD GEMM), bit5 cp.async-prefetching k_coord_ln, bit6 fused edge FFN (edge_ffn_tc.cu: 62 us vs 86 us for the three kernels it replaces), bit7 coordinate update of block l-1 fused into the RBF kernel of block l (k_pos_rbf, per-molecule CTAs; both precision modes), bit8 skip projection edge_l into the edge heads as a third MMA of edge_ffn_kernel (needs bit6)
  // node-chain / edge-chain overlap inside a block (dmt_kernels.cu): the atom-side kernels (18 k rows, one tile per
  // SM, latency-bound) run on a side stream next to the pair-side kernels; persistent GEMM grids are capped so that
  // both fit on the 148 SMs at once.  Opt-in experiment: DS_OVERLAP=1 enables, DS_SPLIT=edge,node sets the caps.
  int overlap = 0;                   // measured on B200: 455.9 ms / 100 steps without, 462-724 ms with (split 100,48 .. 140,8): the atom chain is
                                     // throughput-, not latency-bound, so a branch only takes SMs away from the pair chain
  int edge_cap = 124, node_cap = 24;
  int att_g = 4;                     // DS_ATT_G: targets (= warps) per attention CTA; 4 -> 128-thread CTAs, 12 per SM: 82.1 us vs 84.4 us with 8 (finer tail); 2 targets / 64 threads measured 84.7 us
  int att_p1 = 1;                    // DS_ATT_P1: 1 = source-major pass 1 of the attention (a thread owns (source, head) and walks the targets)
  int tanh_mix = 0;                  // DS_TANH_MIX: 1 = the tanh epilogue of lin_edge0|1 evaluates every second column pair as a polynomial on the FMA
                                     // pipe (common.cuh tanh_poly2) instead of MUFU.TANH.  Measured on B200 ([162 305, 64] x [512, 64]^T, bf16 out):
                                     // no activation 36.8 us, tanh 49.1 us, mix 51.2 us, SiLU 53.4 us, GELU(erf) 125 us: the epilogue is bound by
                                     // issue slots (~3 us per instruction per element), not by MUFU, so the 7 extra instructions per pair lose
  int ffn_variant = 1;               // DS_FFN: 1 = the pre-LayerNorm row is built cooperatively while the tile is staged (55.6 us), 0 = by the row's own thread (64.0 us)
  int cta_cap = 0;                   // cap applied to the next persistent-GEMM launches (0 = all SMs)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  long long launch_count = 0;        // kernels launched (or captured) through this context
  // cached CUDA graph of one sampling step (ds_sample_loop)
  cudaGraphExec_t step_graph = nullptr;
  cudaStream_t capture_stream = nullptr;
  long long step_graph_launches = 0;   // kernels inside one replay of step_graph
  bool step_graph_is_loop = false;     // step_graph holds the WHILE node (whole loop in one launch) rather than one step
  int loop_graph = 0;                  // DS_LOOP_GRAPH: 1 = the whole sampling loop is ONE graph launch (WHILE conditional node whose body is a step);
                                       // measured on B200 at 200 steps: 754.6 ms per round vs 742.5 ms for 200 host launches of the one-step graph
                                       // (+60 us per iteration of the conditional node), hence off by default
};

inline bool ds_is_bf16(const DsContext* c) { return c->mode == 1; }
