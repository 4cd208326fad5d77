// Epilogue <-> global memory for "thread = row" accumulator layouts (tcgen05.ld 32x32b: lane l of a warp holds 32 consecutive columns of
// row l).  Stored straight from that layout, one warp instruction touches 32 rows = 32 different 128-byte lines for 16 useful bytes each:
// the load/store unit spends one cycle per line, and at 96-192 channels (rows 192-384 B apart) those cycles -- not HBM, not the tensor
// pipe -- bounded the vocoder's thin stages (a 128 x 96 fp16 tile: 3 x 1536 line visits against ~2300 MMA cycles).  These helpers turn
// the access around inside the warp, through a 2 KB warp-private shared-memory patch: four lanes cover 64 contiguous bytes of one row,
// one instruction covers 8 rows = 8 lines.  No block-level barrier: a warp only ever touches its own patch.
#pragma once
#include <cstdint>

namespace q3 {
namespace epiio {

constexpr int kPatchBytes = 32 * 64;  // 32 rows x 32 fp16

// 16-byte chunk j (0..3) of row r (0..31) sits at r*64 + ((j ^ ((r >> 1) & 3)) << 4): conflict-free for the row-wise accesses (8 lanes =
// 8 rows, one chunk each) and for the transposed ones (8 lanes = 2 rows x 4 chunks)
__device__ __forceinline__ uint32_t patch_addr(uint32_t patch, int r, int j) { return patch + (uint32_t)(r * 64) + (uint32_t)((j ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ void sts16(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}

// Lane l holds v[0..3] = the 64 bytes (32 fp16) of row l at byte offset `col_bytes` of a [rows][pitch_bytes] global array whose row
// `row0` is this warp's row 0; rows >= rows_valid are not written.  All 32 lanes must call.
__device__ __forceinline__ void warp_store_64B_rows(uint8_t* gbase, size_t row0, size_t pitch_bytes, int col_bytes, const uint4 (&v)[4], uint32_t patch, int lane,
                                                    int rows_valid) {
#pragma unroll
  for (int j = 0; j < 4; ++j) sts16(patch_addr(patch, lane, j), v[j]);
  __syncwarp();
  const int i = lane & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = 8 * k + (lane >> 2);
    const uint4 w = lds16(patch_addr(patch, r, i));
    if (r < rows_valid) *reinterpret_cast<uint4*>(gbase + (row0 + (size_t)r) * pitch_bytes + (size_t)col_bytes + (size_t)(i * 16)) = w;
  }
  __syncwarp();  // the patch may be rewritten
}

// The mirror image: coalesced global reads into the patch; complete() then hands lane l the 64 bytes of row l.  Split in two so that the
// DRAM round trip overlaps other work: issue() only requests the data (4 x 16 B per lane, held in `tmp`), complete() stages and reads back.
__device__ __forceinline__ void warp_load_64B_rows_issue(const uint8_t* gbase, size_t row0, size_t pitch_bytes, int col_bytes, uint4 (&tmp)[4], int lane, int rows_valid) {
  const int i = lane & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = 8 * k + (lane >> 2);
    tmp[k] = r < rows_valid ? *reinterpret_cast<const uint4*>(gbase + (row0 + (size_t)r) * pitch_bytes + (size_t)col_bytes + (size_t)(i * 16)) : make_uint4(0, 0, 0, 0);
  }
}
__device__ __forceinline__ void warp_load_64B_rows_complete(const uint4 (&tmp)[4], uint4 (&v)[4], uint32_t patch, int lane) {
  const int i = lane & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) sts16(patch_addr(patch, 8 * k + (lane >> 2), i), tmp[k]);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = lds16(patch_addr(patch, lane, j));
  __syncwarp();
}

}  // namespace epiio
}  // namespace q3
