// tiled_layout.cuh — global-memory layout of the pre-split GEMM operands of mlp_tc.cuh.
//
// A K-major operand matrix X [rows, Kp] (rows a multiple of 128, Kp a multiple of 32) is stored as
//     [rows / 128][Kp / 32][ 128 x 32 tile ]
// where every 16 KB tile is byte-for-byte the SWIZZLE_128B shared-memory image the tensor core reads: 8-row groups of
// 1024 bytes, 128-byte rows, the 16-byte chunk c of row r at chunk position c ^ (r & 7).  A pipeline stage is then four
// CONTIGUOUS 16 KB reads (one cp.async.bulk each) instead of 512 scattered 128-byte pieces 3.7 KB apart — measured with the
// row-major layout: the copy threads sat on the long scoreboard, 4.5 TB/s L2 -> SM at 30 % LTS utilisation, DRAM pages
// touched 128 bytes at a time.  Everyone who writes such an operand (weight preparation, layer epilogues, the observation
// kernel, split_kernel) goes through offset().
#pragma once
#include <stddef.h>

namespace tiled {

constexpr int kRows = 128, kCols = 32, kTileFloats = kRows * kCols;

#if defined(__CUDACC__)
__host__ __device__
#endif
inline size_t offset(int m, int k, int Kp) {
  const int rt = m >> 7, r = m & 127, kb = k >> 5, kk = k & 31;
  return ((size_t)rt * (size_t)(Kp >> 5) + (size_t)kb) * kTileFloats + (size_t)((r >> 3) * 256 + (r & 7) * 32 + (((kk >> 2) ^ (r & 7)) << 2) + (kk & 3));
}

}  // namespace tiled
