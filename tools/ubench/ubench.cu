// Micro-benchmarks that size the design of the fused STAC solver on sm_100a (numbers land in profiles/ubench_r2.txt).
// Single-warp latencies (clock64 around dependent chains) and whole-chip FP32 throughput for FFMA (3 register operands),
// FFMA with constant operands, and packed FFMA2.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void lat_kernel(float *out, long long *cyc, int iters) {
  const int lane = threadIdx.x & 31;
  float a = 1.0f + lane * 1e-3f, b = 0.999f + lane * 1e-6f, c = 1e-4f;
  __shared__ float sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  long long t0, t1;
  // 0: dependent FFMA chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a = fmaf(a, b, c);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 1: 8 independent FFMA chains, register operands (b varies per lane so not an immediate)
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = a + k;
  float b2 = b * 1.0001f, c2 = c * 1.1f;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = fmaf(v[k], b2, c2);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
#pragma unroll
  for (int k = 0; k < 8; k++) a += v[k];
  // 2: dependent SHFL chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a = __shfl_xor_sync(0xffffffffu, a, 1 + (k & 3));
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // 3: 7 independent shuffles + dependent add (shape of a pose gather)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 7; k++) s += __shfl_sync(0xffffffffu, v[k], (lane + 3) & 31);
      v[0] = s; v[1] = s + 1.f; v[2] = s + 2.f; v[3] = s * 0.5f; v[4] = s - 1.f; v[5] = s - 2.f; v[6] = s * 0.25f;
    }
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  a += v[0];
  // 4: dependent LDS chain (store + load through shared memory, one warp)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      sm[lane] = a;
      __syncwarp();
      a = sm[(lane + 1) & 31];
      __syncwarp();
    }
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 5: __syncthreads with all warps of the block arriving together
  __syncthreads();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) { __syncthreads(); a = a * 1.0001f; }
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 6: dependent FMUL->FADD (cross pipe?) chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) { a = a * b; a = a + c; }
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // 7: dependent FFMA2 chain
  float2 p = make_float2(a, a + 1.f), q = make_float2(b, b), r2 = make_float2(c, c);
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) p = __ffma2_rn(p, q, r2);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  a += p.x + p.y;
  // 8: dependent FSEL / select chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a = (a > b) ? a * 0.5f : a + c;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

template <int MODE>
__global__ void __launch_bounds__(256) tput_kernel(float *out, int iters, float bb, float cc) {
  const float b = bb + threadIdx.x * 1e-7f, c = cc + threadIdx.x * 1e-8f;  // register operands
  float s = 0.f;
  if (MODE == 0) {  // FFMA, 3 register operands, 16 chains
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
  } else if (MODE == 1) {  // FFMA with constants (immediate form)
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], 0.9999f, 1e-4f);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
  } else if (MODE == 2) {  // FFMA2 packed, 8 chains of float2
    float2 a[8];
    const float2 b2 = make_float2(b, b * 1.00001f), c2 = make_float2(c, c * 1.1f);
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(1.0f + 1e-3f * (threadIdx.x + i), 1.0f + 2e-3f * (threadIdx.x + i));
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], b2, c2);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
  } else if (MODE == 3) {  // mix typical of the solver: FFMA + FMUL + FADD with distinct register operands
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        a[i] = fmaf(a[i], a[i + 1], a[i + 2]);
        a[i + 1] = a[i + 1] * b;
        a[i + 2] = a[i + 2] + c;
        a[i + 3] = fmaf(a[i + 3], a[i], c);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
  } else if (MODE == 4) {  // SHFL throughput: 8 independent values, one indexed shuffle each per iteration
    float a[8];
    const int src = (threadIdx.x * 7 + 3) & 31;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 1.0f + 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = __shfl_sync(0xffffffffu, a[i], src);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
  } else if (MODE == 5) {  // the solver's mix: per 8 SHFL, 56 FP32 instructions (12.5 % shuffles)
    float a[8], v[8];
    const int src = (threadIdx.x * 7 + 3) & 31;
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = 1.0f + 1e-3f * (threadIdx.x + i); v[i] = 0.5f + 1e-3f * i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float t = __shfl_sync(0xffffffffu, a[i], src);
        v[i] = fmaf(v[i], b, t);
        v[i] = v[i] * b;
        v[i] = v[i] + c;
        v[i] = fmaf(v[i], b, c);
        v[i] = fmaf(v[i], t, c);
        v[i] = v[i] * t;
        a[i] = fmaf(v[i], b, c);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + v[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run_tput(const char *name, float *out, int sms, double flop_per_thread_iter, int blocks_per_sm = 8, double instr_per_thread_iter = 0, int clk_khz = 0) {
  const int blocks = sms * blocks_per_sm, threads = 256, iters = 20000;
  tput_kernel<MODE><<<blocks, threads>>>(out, 200, 0.9999f, 1e-4f);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    tput_kernel<MODE><<<blocks, threads>>>(out, iters, 0.9999f, 1e-4f);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  if (instr_per_thread_iter > 0)  // warp instructions per clock per SM
    printf("tput %-28s %8.3f warp-instr/clk/SM at %d warps/SM (%.3f ms)\n", name,
           instr_per_thread_iter * iters * (double)blocks * threads / 32.0 / (best * 1e-3) / (clk_khz * 1e3) / sms, blocks_per_sm * threads / 32, best);
  else
    printf("tput %-28s %8.2f TFLOP/s  (%.3f ms)\n", name, flop_per_thread_iter * iters * (double)blocks * threads / (best * 1e-3) / 1e12, best);
}

int main() {
  int dev = 0, sms = 0, clk = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  printf("SMs %d, clock %d kHz\n", sms, clk);
  float *out; long long *cyc;
  CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256));
  CK(cudaMallocManaged(&cyc, sizeof(long long) * 16));
  const int iters = 2000;
  const char *names[] = {"dependent FFMA", "8 indep FFMA chains (per instr)", "dependent SHFL", "7 SHFL + add tree (per group)", "STS+syncwarp+LDS+syncwarp round trip",
                         "__syncthreads (+FMUL)", "dependent FMUL->FADD (per instr)", "dependent FFMA2", "dependent compare+select+op"};
  const double per[] = {16, 16, 16, 2, 16, 16, 16, 16, 16};
  for (int warps = 1; warps <= 6; warps += (warps == 1 ? 3 : 2)) {
    lat_kernel<<<1, 32 * warps>>>(out, cyc, 10);
    CK(cudaDeviceSynchronize());
    lat_kernel<<<1, 32 * warps>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    printf("-- block of %d warp(s): cycles per item (warp 0)\n", warps);
    for (int i = 0; i < 9; i++) printf("lat %-40s %8.2f\n", names[i], (double)cyc[i] / (iters * per[i]));
  }
  run_tput<0>("FFMA reg,reg,reg", out, sms, 32.0);
  run_tput<1>("FFMA reg,imm,imm", out, sms, 32.0);
  run_tput<2>("FFMA2 packed", out, sms, 32.0);
  run_tput<3>("FFMA/FMUL/FADD mix", out, sms, 24.0);  // 4 groups x (2 fma + mul + add) = 24 flop
  for (int bps = 1; bps <= 8; bps *= 2) run_tput<4>("SHFL (indexed)", out, sms, 0, bps, 8.0, clk);
  for (int bps = 1; bps <= 2; bps *= 2) run_tput<5>("8 SHFL + 56 FP32 (all instr)", out, sms, 0, bps, 64.0, clk);
  return 0;
}
