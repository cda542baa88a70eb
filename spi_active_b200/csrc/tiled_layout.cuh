// tiled_layout.cuh — global-memory layout of the pre-split GEMM operands of mlp_tc.cuh.
//
// Operands are fp16 PAIRS (x * scale = hi + lo, both fp16: the "3 x FP16" split, see mlp_tc.cuh).  A K-major operand matrix
// X [rows, Kp] (rows a multiple of 128, Kp a multiple of 64) is stored as
//     [rows / 128][Kp / 64][ 128 x 64 tile ]
// where every 16 KB tile is byte-for-byte the SWIZZLE_128B shared-memory image the tensor core reads: 8-row groups of
// 1024 bytes, 128-byte rows (64 halves of K), the 16-byte chunk c of row r at chunk position c ^ (r & 7).  A pipeline stage is
// then four CONTIGUOUS 16 KB reads (one cp.async.bulk each) instead of 512 scattered 128-byte pieces — measured with a
// row-major layout: the copy threads sat on the long scoreboard, DRAM pages touched 128 bytes at a time.  Everyone who
// writes such an operand (weight preparation, layer epilogues, the observation kernel, split_kernel) goes through offset().
#pragma once
#include <stddef.h>

namespace tiled {

constexpr int kRows = 128, kCols = 64, kTileElems = kRows * kCols;   // elements (halves) per 16 KB tile

#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t offset(int m, int k, int Kp) {     // element (half) index of X[m][k]
  const int rt = m >> 7, r = m & 127, kb = k >> 6, kk = k & 63;
  return ((size_t)rt * (size_t)(Kp >> 6) + (size_t)kb) * kTileElems +
         (size_t)((r >> 3) * 512 + (r & 7) * 64 + (((kk >> 3) ^ (r & 7)) << 3) + (kk & 7));
}

}  // namespace tiled
