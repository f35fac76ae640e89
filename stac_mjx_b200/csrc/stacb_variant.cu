// One kernel variant per translation unit so the variants compile in parallel.
// Built with -DV_CPL=.. -DV_NB=.. -DV_NBF=.. -DV_SPL=.. -DV_JM=.. (see build.sh / stacb_variants.h).
#include "stacb_device.cuh"
#include "stacb_variants.h"

namespace stacb {

#define FN_(prefix, a, b, c, d, e) prefix##a##_##b##_##c##_##d##_##e
#define FN(prefix, a, b, c, d, e) FN_(prefix, a, b, c, d, e)

cudaError_t FN(launch_pose_, V_CPL, V_NB, V_NBF, V_SPL, V_JM)(const DevTree &T, const PoseArgs &a, int grid, int block, size_t smem, int coop,
                                                        cudaStream_t s) {
  // grouped latency mode (3): each of the GRP member warps of a role carries ceil(V_NB / GRP) bodies per lane (at least 2, so
  // the shared-memory pose exchange is used); only instantiated for variants with more than one body per lane
  constexpr int NBG = (V_NB + GRP - 1) / GRP < 2 ? 2 : (V_NB + GRP - 1) / GRP;
  auto k = coop == 1 ? pose_clips_kernel<V_CPL, V_NB, V_NBF, V_SPL, 1>
         : coop == 2 ? pose_clips_kernel<V_CPL, V_NB, V_NBF, V_SPL, 2>
         : (coop == 3 && V_NB > 1) ? pose_clips_kernel<V_CPL, (V_NB > 1 ? NBG : V_NB), V_NBF, V_SPL, (V_NB > 1 ? 3 : 0)>
                                   : pose_clips_kernel<V_CPL, V_NB, V_NBF, V_SPL, 0>;
  if (smem > 40 * 1024) {  // dynamic + static (momentum table, counters: ~2 KB) beyond the 48 KB default needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, s>>>(T, a);
  return cudaGetLastError();
}

cudaError_t FN(launch_batch_, V_CPL, V_NB, V_NBF, V_SPL, V_JM)(const DevTree &T, const BatchArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
  auto k = batch_kernel<V_CPL, V_NB, V_NBF, V_SPL>;
  if (smem > 40 * 1024) {  // dynamic + static (momentum table, counters: ~2 KB) beyond the 48 KB default needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, s>>>(T, a);
  return cudaGetLastError();
}

cudaError_t FN(launch_mphase_, V_CPL, V_NB, V_NBF, V_SPL, V_JM)(const DevTree &T, const MArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
  auto k = m_phase_kernel<V_CPL, V_NB, V_NBF, V_SPL>;
  if (smem > 40 * 1024) {  // dynamic + static (momentum table, counters: ~2 KB) beyond the 48 KB default needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, s>>>(T, a);
  return cudaGetLastError();
}

}  // namespace stacb
