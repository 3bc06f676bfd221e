// blcd_peaks.cu -- measured denominators for the solver's roofline (include/boxlcd_b200.h: blcd_measure_peaks).
//
// The rollout kernel is bound by instruction issue / dependent fp32 latency, not by HBM, and
// MEASURED_PEAKS.json carries no fp32 figure.  This file measures, on the device the bench runs on and inside the bench
// run, the two rates that bound it:
//   * fp32 FMA throughput: every lane retires one FFMA per issue slot -- 8 independent accumulator chains per thread,
//     16 resident warps per scheduler, so neither the 4-cycle dependent latency nor issue gaps limit it.  One FFMA is one
//     lane-instruction, so the same number is the peak LANE-INSTRUCTION rate (warp issue slots x 32 lanes);
//   * the dependent-chain rate: ONE accumulator chain per thread and one warp per scheduler, i.e. how fast a purely serial
//     fp32 recurrence advances -- the regime a Gauss-Seidel sweep of one world is in.
#include <cuda_runtime.h>
#include <stdint.h>
#include "boxlcd_b200.h"

extern "C" int blcd_fail_msg(const char* msg);

namespace {

template <int CHAINS>
__global__ void __launch_bounds__(512) k_fma_peak(float* out, int iters, float a, float b) {
  float acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = (float)(threadIdx.x + c);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) acc[c] = fmaf(acc[c], a, b);
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the loop alive
}

template <int CHAINS>
int time_fma(int blocks, int threads, int iters, double* lane_fma_per_s) {
  float* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * 4) != cudaSuccess) return blcd_fail_msg("blcd_measure_peaks: cudaMalloc failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k_fma_peak<CHAINS><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess || best > 1e29f) return blcd_fail_msg("blcd_measure_peaks: kernel failed");
  *lane_fma_per_s = (double)blocks * threads * (double)iters * 8.0 * CHAINS / (best * 1e-3);
  return 0;
}

}  // namespace

extern "C" int blcd_measure_peaks(int device, double* out4) {
  if (!out4) return blcd_fail_msg("blcd_measure_peaks: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return blcd_fail_msg("blcd_measure_peaks: no such CUDA device");
  cudaSetDevice(device);
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  double full = 0.0, chain = 0.0;
  // 4 blocks of 512 threads per SM = 64 warps per SM (16 per scheduler), 8 chains each
  if (time_fma<8>(sms * 4, 512, 4096, &full)) return -1;
  // 1 block of 128 threads per SM = one warp per scheduler, one chain: lanes x (1 FFMA per dependent latency)
  if (time_fma<1>(sms, 128, 1 << 16, &chain)) return -1;
  out4[0] = full;                       // lane-FMAs per second, whole device (x2 = FLOP/s)
  out4[1] = chain / ((double)sms * 128);// FFMAs per second of ONE dependent chain (clock / dependent latency)
  out4[2] = (double)sms;
  out4[3] = (double)khz * 1e3;          // nominal max SM clock, Hz
  return 0;
}
