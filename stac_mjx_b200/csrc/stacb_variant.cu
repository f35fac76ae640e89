// One kernel variant per translation unit so the variants compile in parallel.
// Built with -DV_CPL=.. -DV_NB=.. -DV_NBF=.. -DV_SPL=.. (see build.sh / stacb_variants.h).
#include "stacb_device.cuh"
#include "stacb_variants.h"

namespace stacb {

#define CAT_(a, b, c, d) a##_##b##_##c##_##d
#define FN_(prefix, a, b, c, d) prefix##a##_##b##_##c##_##d
#define FN(prefix, a, b, c, d) FN_(prefix, a, b, c, d)

cudaError_t FN(launch_pose_, V_CPL, V_NB, V_NBF, V_SPL)(const DevTree &T, const PoseArgs &a, int grid, int block, size_t smem, int coop,
                                                        cudaStream_t s) {
  auto k = coop ? pose_clips_kernel<V_CPL, V_NB, V_NBF, V_SPL, true> : pose_clips_kernel<V_CPL, V_NB, V_NBF, V_SPL, false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, s>>>(T, a);
  return cudaGetLastError();
}

cudaError_t FN(launch_batch_, V_CPL, V_NB, V_NBF, V_SPL)(const DevTree &T, const BatchArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
  auto k = batch_kernel<V_CPL, V_NB, V_NBF, V_SPL>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, s>>>(T, a);
  return cudaGetLastError();
}

}  // namespace stacb
