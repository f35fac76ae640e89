// One kernel variant per translation unit so the variants compile in parallel.
// Built with -DV_CPL=.. -DV_NB=.. -DV_NBF=.. -DV_SPL=.. -DV_JM=.. (see build.sh / stacb_variants.h).
#include "stacb_device.cuh"
#include "stacb_variants.h"

namespace stacb {

#define FN_(prefix, a, b, c, d, e) prefix##a##_##b##_##c##_##d##_##e
#define FN(prefix, a, b, c, d, e) FN_(prefix, a, b, c, d, e)

cudaError_t FN(launch_pose_, V_CPL, V_NB, V_NBF, V_SPL, V_JM)(const DevTree &T, const PoseArgs &a, int grid, int block, size_t smem, int coop,
                                                        cudaStream_t s) {
  auto k = coop == 1 ? pose_clips_kernel<V_CPL, V_NB, V_NBF, V_SPL, 1>
         : coop == 2 ? pose_clips_kernel<V_CPL, V_NB, V_NBF, V_SPL, 2> : pose_clips_kernel<V_CPL, V_NB, V_NBF, V_SPL, 0>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, s>>>(T, a);
  return cudaGetLastError();
}

cudaError_t FN(launch_batch_, V_CPL, V_NB, V_NBF, V_SPL, V_JM)(const DevTree &T, const BatchArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
  auto k = batch_kernel<V_CPL, V_NB, V_NBF, V_SPL>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, s>>>(T, a);
  return cudaGetLastError();
}

}  // namespace stacb
