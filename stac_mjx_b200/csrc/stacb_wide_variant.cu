// The multi-warp register-resident solver (stacb_wide.cuh), one translation unit per warp count.  Built with -DV_WW=2|4|6|8.
#include "stacb_wide.cuh"

namespace stacb {

#define FN_(prefix, a) prefix##a
#define FN(prefix, a) FN_(prefix, a)

// sched: 0 one group of W warps per chain, 1 the same with registers capped for two CTAs per SM (W <= 6), 2 pair mode (two groups
// of W warps per chain, W <= 8: 512 threads)
cudaError_t FN(launch_wide_pose_, V_WW)(const DevTree &T, const PoseArgs &a, int grid, size_t area_bytes, int sched, cudaStream_t s) {
  auto k = wide::wide_pose_kernel<V_WW, 8, 1, 1>;
  int groups = 1;
  if (sched == 1 && V_WW <= 6) k = wide::wide_pose_kernel<V_WW, 8, (V_WW <= 6 ? 2 : 1), 1>;
  if (sched == 2) { k = wide::wide_pose_kernel<V_WW, 8, 1, 2>; groups = 2; }
  const size_t smem = area_bytes + 16 + groups * sizeof(wide::WX<V_WW>) + (groups == 2 ? sizeof(wide::XPW<V_WW>) : 0);  // + alignment slack
  if (smem > 40 * 1024) {  // dynamic + static (momentum table, counters: ~2 KB) beyond the 48 KB default needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, 32 * V_WW * groups, smem, s>>>(T, a);
  return cudaGetLastError();
}

cudaError_t FN(launch_wide_batch_, V_WW)(const DevTree &T, const BatchArgs &a, int grid, cudaStream_t s) {
  wide::wide_batch_kernel<V_WW><<<grid, 32 * V_WW, sizeof(wide::WX<V_WW>), s>>>(T, a);
  return cudaGetLastError();
}

}  // namespace stacb
