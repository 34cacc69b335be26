// TMA (cp.async.bulk.tensor) plumbing for the sm_100a kernels: host-side tensor-map encoding through the driver entry
// point (no link against libcuda) and the device-side PTX wrappers.
//
// Layout facts the kernels rely on (probed on a B200 with tools/b200_probe.cu, see profiles/):
//  * fp32 boxes of [rows][32 columns] (128-byte rows) land in shared memory as [rows][128 B];
//  * CU_TENSOR_MAP_SWIZZLE_128B XORs the 16-byte chunk index with (row & 7)  == tc::sw128_off: the K-major UMMA operand layout;
//  * CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B XORs the 32-byte chunk index with (row & 3) == tc::mn32_off: the MN-major
//    (SWIZZLE_128B_BASE32B) UMMA operand layout, so a row-major [batch][features] matrix is an MN-major operand as it lies;
//  * elements outside the tensor are written as zeros and still count towards the transaction bytes.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bbb_common.cuh"

namespace bbb {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn();   // bbb_mlp.cu: resolved once through cudaGetDriverEntryPointByVersion

enum Swz { kNone = 0, kSw128 = 1, kSw128Atom32 = 2 };

// fp32 tensor [d2][d1][d0] (d0 contiguous; d2 == 0: rank 2) with a box of [1][box1][box0]; d0 * 4 and the base must be
// multiples of 16 bytes.  Returns BBB_OK or sets the error string.
inline int make_map(CUtensorMap *map, const float *base, int64_t d0, int64_t d1, int64_t d2, int box0, int box1, Swz swz) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(BBB_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const int rank = d2 > 0 ? 3 : 2;
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)(d2 > 0 ? d2 : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * 4, (cuuint64_t)d0 * (cuuint64_t)d1 * 4};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1}, es[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swz == kNone ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : swz == kSw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float *>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(BBB_ECUDA, "cuTensorMapEncodeTiled failed (%d) for dims (%lld, %lld, %lld) box (%d, %d)", (int)r,
                (long long)d0, (long long)d1, (long long)d2, box0, box1);
  return BBB_OK;
}

// General form: `rank` dims (dims[0] contiguous), byte strides of dims 1.. (multiples of 16), box extents.
inline int make_map_nd(CUtensorMap *map, const float *base, int rank, const int64_t *dims_in, const int64_t *strides_in,
                       const int *box_in, Swz swz) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(BBB_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], es[5];
  for (int i = 0; i < rank; ++i) { dims[i] = (cuuint64_t)dims_in[i]; box[i] = (cuuint32_t)box_in[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) strides[i] = (cuuint64_t)strides_in[i];
  const CUtensorMapSwizzle sw = swz == kNone ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : swz == kSw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float *>(base), dims, strides,
                         box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BBB_ECUDA, "cuTensorMapEncodeTiled failed (%d) for a rank-%d map", (int)r, rank);
  return BBB_OK;
}
// A row-major [S][rows][cols] matrix (cols % 32 == 0) seen as [S][cols / 32 groups][rows][32]: one box of
// [1][groups][box_rows][32] lands as `groups` stacked [box_rows][128 B] regions -- with kSw128Atom32 the MN-major UMMA
// operand slab of `groups * 32` columns, with kSw128 `groups` K-major tiles of 32 k each.
inline int make_map_grouped(CUtensorMap *map, const float *base, int64_t cols, int64_t rows, int64_t S, int box_rows,
                            int groups, Swz swz) {
  const int64_t dims[4] = {32, rows, cols / 32, S > 0 ? S : 1};
  const int64_t strides[3] = {cols * 4, 128, rows * cols * 4};
  const int box[4] = {32, box_rows, groups, 1};
  return make_map_nd(map, base, 4, dims, strides, box, swz);
}

// ---- device side ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_map(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// arrive on `bar` and announce `bytes` of TMA traffic that will complete on it
__device__ __forceinline__ void arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void load_2d(uint32_t smem_dst, const CUtensorMap *m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void load_3d(uint32_t smem_dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void load_4d(uint32_t smem_dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// shared -> global reduction: every element of the [1][box1][box0] tile at `smem_src` is ADDED to the tensor (fp32 add
// performed in L2; elements outside the tensor are dropped).  Bulk-group completion: commit, then wait.
__device__ __forceinline__ void reduce_add_3d(const CUtensorMap *m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// until the committed bulk operations have READ their shared-memory source (which may then be reused or released); their
// global writes complete by the end of the grid at the latest, which is what a dependent grid waits for
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

}  // namespace tma
}  // namespace bbb
