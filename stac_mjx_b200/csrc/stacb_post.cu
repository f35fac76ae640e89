// stacb_post.cu -- device epilogues of the IK pass, run on the packed outputs before they leave the GPU:
//   stacb_edge_crossfade : reference stac_mjx/utils.py:393-461 (handle_edge_effects: sigmoid cross-fade of the look-ahead overlap of
//                          `continuous` clips, overlaps removed)
//   stacb_qvel           : reference stac_mjx/utils.py:302-347 (compute_velocity_from_kinematics) with its quaternion helpers :195-299
// Both are pure streaming kernels (one read, one write of each element): HBM-bound, coalesced along the feature dimension.
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "stacb.h"

namespace stacb {

int post_fail(int code, const char *msg);  // stacb_abi.cu: records the thread-local error message

// out row r -> (clip, frame) of the [C, F + ov, D] input; rows of clip c < C - 1 with frame >= F are blended with the first ov frames
// of clip c + 1.  The blend is evaluated in float64 and rounded to float32, exactly as numpy does for (1.0 - m) * a + m * b with a
// float64 weight vector and float32 data.
__global__ void edge_crossfade_kernel(const float *__restrict__ in, const double *__restrict__ w, float *__restrict__ out, int C, int F, int ov,
                                      int D, long long n_rows) {
  const long long total = n_rows * D;
  const int FE = F + ov;
  const long long mid = (long long)(C > 2 ? C - 2 : 0) * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D;
    const int d = (int)(i - r * D);
    int c, f;
    if (r < FE) { c = 0; f = (int)r; }
    else if (r < FE + mid) { const long long k = r - FE; c = 1 + (int)(k / F); f = ov + (int)(k % F); }
    else { c = C - 1; f = ov + (int)(r - FE - mid); }
    const float a = in[((long long)c * FE + f) * D + d];
    float v = a;
    if (c < C - 1 && f >= F) {
      const int j = f - F;
      const float b = in[((long long)(c + 1) * FE + j) * D + d];
      const double m = w[j];
      v = (float)((1.0 - m) * (double)a + m * (double)b);
    }
    out[i] = v;
  }
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// one thread per (clip, frame); the last frame of a clip is differenced against itself (zero velocity), as the reference's padding does
__global__ void qvel_kernel(const float *__restrict__ qpos, float *__restrict__ qvel, int C, int F, int nq, int freejoint, float dt, float max_qvel) {
  const long long n = (long long)C * F;
  const int nv = freejoint ? nq - 1 : nq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % F);
    const float *q0 = qpos + i * nq;
    const float *q1 = (t + 1 < F) ? q0 + nq : q0;
    float *o = qvel + i * nv;
    if (!freejoint) {
      for (int k = 0; k < nq; k++) o[k] = clampf((q1[k] - q0[k]) / dt, -max_qvel, max_qvel);
      continue;
    }
    for (int k = 0; k < 3; k++) o[k] = (q1[k] - q0[k]) / dt;
    // quat_diff(source, target) = conj(source) * target, normalised, then quat_to_axisangle
    const float aw = q0[3], ax = -q0[4], ay = -q0[5], az = -q0[6];
    const float bw = q1[3], bx = q1[4], by = q1[5], bz = q1[6];
    float dw = aw * bw - ax * bx - ay * by - az * bz;
    float dx = aw * bx + ax * bw + ay * bz - az * by;
    float dy = aw * by - ax * bz + ay * bw + az * bx;
    float dz = aw * bz + ax * by - ay * bx + az * bw;
    const float nrm = sqrtf(dw * dw + dx * dx + dy * dy + dz * dz);
    dw /= nrm; dx /= nrm; dy /= nrm; dz /= nrm;
    const float angle = 2.0f * acosf(clampf(dw, -1.0f, 1.0f));
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (!(angle < 1e-10f)) {
      const float qn = sinf(angle / 2.0f);
      const float pi = 3.14159274f;
      float wrapped = fmodf(angle + pi, 2.0f * pi);
      if (wrapped < 0.f) wrapped += 2.0f * pi;
      wrapped -= pi;
      gx = dx / qn * wrapped; gy = dy / qn * wrapped; gz = dz / qn * wrapped;
    }
    o[3] = gx / dt; o[4] = gy / dt; o[5] = gz / dt;
    for (int k = 7; k < nq; k++) o[k - 1] = clampf((q1[k] - q0[k]) / dt, -max_qvel, max_qvel);
  }
}

}  // namespace stacb

using namespace stacb;

extern "C" long long stacb_edge_rows(int C, int F, int ov) {
  if (C < 1 || F < ov || ov < 0) return -1;
  return (long long)(F + ov) + (long long)(C > 2 ? C - 2 : 0) * F + (F - ov);
}

extern "C" int stacb_edge_crossfade(const float *in, const double *w, float *out, int C, int F, int ov, int D, void *stream) {
  if (!in || (!w && ov > 0) || !out || C < 1 || D < 1 || ov < 0 || F < ov) return post_fail(STACB_E_INVALID, "stacb_edge_crossfade: bad argument");
  const long long rows = stacb_edge_rows(C, F, ov), total = rows * D;
  const int block = 256;
  const int grid = (int)std::min<long long>((total + block - 1) / block, 148LL * 32);
  edge_crossfade_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, w, out, C, F, ov, D, rows);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? STACB_OK : post_fail(STACB_E_CUDA, cudaGetErrorString(e));
}

extern "C" int stacb_qvel(const float *qpos, float *qvel, int C, int F, int nq, int freejoint, float dt, float max_qvel, void *stream) {
  if (!qpos || !qvel || C < 0 || F < 1 || nq < (freejoint ? 7 : 1) || !(dt > 0.f)) return post_fail(STACB_E_INVALID, "stacb_qvel: bad argument");
  if (C == 0) return STACB_OK;
  const long long n = (long long)C * F;
  const int block = 128;
  const int grid = (int)std::min<long long>((n + block - 1) / block, 148LL * 32);
  qvel_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(qpos, qvel, C, F, nq, freejoint, dt, max_qvel);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? STACB_OK : post_fail(STACB_E_CUDA, cudaGetErrorString(e));
}
