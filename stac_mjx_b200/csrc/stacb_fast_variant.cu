// One variant of the register-resident hinge-tree solver per translation unit (stacb_fast.cuh).
// Built with -DV_FJM=.. -DV_FRT=.. -DV_FNBF=.. (joint slots per body, pointer-jumping rounds, all bodies per lane).
#include "stacb_fast.cuh"
#include "stacb_variants.h"

namespace stacb {

#define FN_(prefix, a, b, c) prefix##a##_##b##_##c
#define FN(prefix, a, b, c) FN_(prefix, a, b, c)

template <class K>
static cudaError_t launch(K k, const DevTree &T, const PoseArgs &a, int grid, int block, size_t smem, cudaStream_t s) {
  if (smem > 40 * 1024) {  // dynamic + static (momentum table, counters: ~2 KB) beyond the 48 KB default needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, s>>>(T, a);
  return cudaGetLastError();
}

// sched: 0 throughput, 1 latency with 4 warps per chain, 2 dense throughput, 3 latency with 6 warps per chain, 4 pair mode (2 warps per chain)
cudaError_t FN(launch_fast_pose_, V_FJM, V_FRT, V_FNBF)(const DevTree &T, const PoseArgs &a, int grid, int block, size_t smem, int sched, cudaStream_t s) {
  switch (sched) {
    case 1: return launch(fast::fast_pose_kernel<V_FJM, V_FRT, V_FNBF, 2, 1>, T, a, grid, block, smem, s);
    case 3: return launch(fast::fast_pose_kernel<V_FJM, V_FRT, V_FNBF, 3, 1>, T, a, grid, block, smem, s);
    case 2: return launch(fast::fast_pose_kernel<V_FJM, V_FRT, V_FNBF, 0, 4>, T, a, grid, block, smem, s);
    case 4: return launch(fast::fast_pose_kernel<V_FJM, V_FRT, V_FNBF, 1, 1>, T, a, grid, block, smem, s);
    default: return launch(fast::fast_pose_kernel<V_FJM, V_FRT, V_FNBF, 0, 1>, T, a, grid, block, smem, s);
  }
}

cudaError_t FN(launch_fast_batch_, V_FJM, V_FRT, V_FNBF)(const DevTree &T, const BatchArgs &a, int grid, int block, cudaStream_t s) {
  fast::fast_batch_kernel<V_FJM, V_FRT><<<grid, block, 0, s>>>(T, a);
  return cudaGetLastError();
}

}  // namespace stacb
