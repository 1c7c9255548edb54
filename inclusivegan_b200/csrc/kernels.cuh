// Device code of libb200knn: the kernels of the IMLE matching path, one header per kernel family.
//
//   convert.cuh   convert_norm_kernel   HBM-bound.  rows of f64/f32 -> centred BF16 rows (TMA-friendly pitch) + ||x~||^2
//                 (fp32, of the rounded values) + ||x - x~||, the exact rounding perturbation (for the certificate).
//   dist.cuh      dist_topc_kernel      tensor-bound.  Q x N scores  s~ = ||x~||^2 - 2 q~.x~  as a BF16 GEMM on tcgen05
//                 (TMA -> smem ring -> tcgen05.mma, fp32 accumulators double-buffered in TMEM) with a fused per-row
//                 top-C selection (or threshold collection) in the epilogue: the Q x N matrix never reaches HBM.
//   rerank.cuh    rerank_kernel         merges the per-stream shortlists, recomputes the survivors exactly (float64
//                 accumulation of (q-x)^2 over the ORIGINAL f64/f32 rows, the arithmetic of the reference's
//                 compute_dist, dci_code/src/util.c:62-69), selects k, and CERTIFIES the answer: a query is exact when
//                 its k-th exact distance is below a rigorous lower bound on the true distance of every row the BF16
//                 pass dropped.  rerank_collect_kernel: the same for the second pass's collected lists.
//   scan.cuh      scan_*                exact float64 CUDA-core scan: overflowed second-pass lists and k > 32.
//   member.cuh    ball_*                ball membership for the k-NN precision/recall metric.
//   exchange.cuh  merge_topk_kernel, publish_topk_kernel, merge_wait_kernel: multi-GPU row sharding.
//   project.cuh   project_kernel        the trainer's optional random projection (float64 GEMM).
#pragma once
#include "common.cuh"
#include "convert.cuh"
#include "tiers.cuh"
#include "dist.cuh"
#include "rerank.cuh"
#include "scan.cuh"
#include "member.cuh"
#include "exchange.cuh"
#include "project.cuh"
