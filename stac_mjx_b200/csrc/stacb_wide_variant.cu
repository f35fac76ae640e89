// The multi-warp register-resident solver (stacb_wide.cuh), one translation unit per warp count.  Built with -DV_WW=2|4|6|8.
#include "stacb_wide.cuh"

namespace stacb {

#define FN_(prefix, a) prefix##a
#define FN(prefix, a) FN_(prefix, a)

cudaError_t FN(launch_wide_pose_, V_WW)(const DevTree &T, const PoseArgs &a, int grid, size_t area_bytes, int dense, cudaStream_t s) {
  auto k = (dense && V_WW <= 6) ? wide::wide_pose_kernel<V_WW, 8, (V_WW <= 6 ? 2 : 1)> : wide::wide_pose_kernel<V_WW, 8, 1>;
  const size_t smem = area_bytes + sizeof(wide::WX<V_WW>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, 32 * V_WW, smem, s>>>(T, a);
  return cudaGetLastError();
}

cudaError_t FN(launch_wide_batch_, V_WW)(const DevTree &T, const BatchArgs &a, int grid, cudaStream_t s) {
  wide::wide_batch_kernel<V_WW><<<grid, 32 * V_WW, sizeof(wide::WX<V_WW>), s>>>(T, a);
  return cudaGetLastError();
}

}  // namespace stacb
