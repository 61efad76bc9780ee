// vecnorm.cu — VecNormalize on the device (SB3 1.0 VecNormalize as built at reference drloco/common/utils.py:130-132).
//
// Two kernels per step, with the (optional) NCCL all-reduce of the packed moments between them:
//   moments: ret = ret*gamma + rew;  packed = [ sum_obs[D], sumsq_obs[D], n, sum_ret, sumsq_ret ]   (float64 sums)
//   apply  : Chan merge of the batch moments into the running statistics (identical on every rank after the
//            all-reduce), obs <- clip((obs-mean)/sqrt(var+eps)), rew <- clip(rew/sqrt(ret_var+eps)), ret[done] = 0
// Running statistics layout (float64): rms = [ mean[D], var[D], count, ret_mean, ret_var, ret_count ].
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/drloco_b200.h"

namespace drl {

constexpr int kVnThreads = 256;

// grid-stride over envs; each thread owns one obs column (t % D) of a strip of rows so that loads stay coalesced.
__global__ void vecnorm_moments_kernel(const float* __restrict__ obs, int n, int d, const float* __restrict__ rew,
                                       float* __restrict__ ret, float gamma, double* __restrict__ packed) {
  extern __shared__ double sh[];      // [2*d + 2]
  for (int i = threadIdx.x; i < 2 * d + 2; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const long long total = (long long)n * d;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // obs: element index e -> column e % d.  Use a stride that is a multiple of d so each thread keeps one column.
  const long long tpb = (long long)blockDim.x * gridDim.x;
  const long long step = (tpb / d) * d;
  if (step > 0) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < step) {
      double s = 0.0, ss = 0.0;
      for (long long e = t; e < total; e += step) {
        const double x = (double)obs[e];
        s += x; ss += x * x;
      }
      const int col = (int)(t % d);
      atomicAdd(&sh[col], s);
      atomicAdd(&sh[d + col], ss);
    }
  }
  if (rew != nullptr) {
    double s = 0.0, ss = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float r = ret[i] * gamma + rew[i];
      ret[i] = r;
      s += (double)r; ss += (double)r * (double)r;
    }
    // warp reduce before the shared atomics
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[2 * d], s); atomicAdd(&sh[2 * d + 1], ss); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) atomicAdd(&packed[i], sh[i]);
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) atomicAdd(&packed[2 * d], (double)n);
    atomicAdd(&packed[2 * d + 1], sh[2 * d]);
    atomicAdd(&packed[2 * d + 2], sh[2 * d + 1]);
  }
}

__device__ __forceinline__ void chan_merge(double mean, double var, double count, double bsum, double bsumsq,
                                           double bn, double& nmean, double& nvar, double& ncount) {
  // RunningMeanStd.update_from_moments (SB3 common/running_mean_std.py)
  const double bmean = bsum / bn;
  double bvar = bsumsq / bn - bmean * bmean;
  if (bvar < 0.0) bvar = 0.0;
  const double delta = bmean - mean, tot = count + bn;
  nmean = mean + delta * bn / tot;
  const double m2 = var * count + bvar * bn + delta * delta * count * bn / tot;
  nvar = m2 / tot;
  ncount = tot;
}

__global__ void vecnorm_apply_kernel(const float* __restrict__ obs_in, float* __restrict__ obs_out,
                                     const float* __restrict__ rew_in, float* __restrict__ rew_out, int n, int d,
                                     const double* __restrict__ packed, const double* __restrict__ rms_in,
                                     double* __restrict__ rms_out, float* __restrict__ ret,
                                     const unsigned char* __restrict__ done, float clip_obs, float clip_rew, float eps,
                                     int training, int norm_obs, int norm_rew) {
  extern __shared__ float shf[];     // mean[d], inv_std[d], ret_inv_std
  float* s_mean = shf;
  float* s_inv = shf + d;
  for (int k = threadIdx.x; k <= d; k += blockDim.x) {
    double mean, var, count;
    if (k < d) { mean = rms_in[k]; var = rms_in[d + k]; count = rms_in[2 * d]; }
    else { mean = rms_in[2 * d + 1]; var = rms_in[2 * d + 2]; count = rms_in[2 * d + 3]; }
    const double bn = packed ? packed[2 * d] : 0.0;
    if (training && bn > 0.0) {
      const double bs = k < d ? packed[k] : packed[2 * d + 1];
      const double bss = k < d ? packed[d + k] : packed[2 * d + 2];
      if (k < d || rew_in != nullptr) chan_merge(mean, var, count, bs, bss, bn, mean, var, count);
    }
    if (blockIdx.x == 0) {
      if (k < d) { rms_out[k] = mean; rms_out[d + k] = var; if (k == 0) rms_out[2 * d] = count; }
      else { rms_out[2 * d + 1] = mean; rms_out[2 * d + 2] = var; rms_out[2 * d + 3] = count; }
    }
    if (k < d) { s_mean[k] = (float)mean; s_inv[k] = (float)(1.0 / sqrt(var + (double)eps)); }
    else s_inv[d] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const long long total = (long long)n * d;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int col = (int)(e % d);
    float x = obs_in[e];
    if (norm_obs) x = fminf(fmaxf((x - s_mean[col]) * s_inv[col], -clip_obs), clip_obs);
    obs_out[e] = x;
  }
  if (rew_in != nullptr) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float r = rew_in[i];
      if (norm_rew) r = fminf(fmaxf(r * s_inv[d], -clip_rew), clip_rew);
      rew_out[i] = r;
      if (done != nullptr && done[i]) ret[i] = 0.f;
    }
  }
}

}  // namespace drl

static thread_local char g_verr[256] = "";
extern "C" const char* drl_vecnorm_last_error(void) { return g_verr; }

// obs [n][d] device, rew [n] device (nullable: reset-time update of the observation statistics only), ret [n] device
// in/out, packed device float64 [2d+3] (zeroed here, then filled).
extern "C" int drl_vecnorm_moments(const float* obs, int32_t n, int32_t d, const float* rew, float* ret, float gamma,
                                   double* packed, void* stream) {
  if (!obs || !packed || n <= 0 || d <= 0 || (rew && !ret)) return DRL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(packed, 0, (2 * d + 3) * sizeof(double), st) != cudaSuccess) return DRL_ERR_CUDA;
  long long work = (long long)n * d;
  int blocks = (int)((work + drl::kVnThreads * 8 - 1) / (drl::kVnThreads * 8));
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;
  drl::vecnorm_moments_kernel<<<blocks, drl::kVnThreads, (2 * d + 2) * sizeof(double), st>>>(obs, n, d, rew, ret, gamma,
                                                                                             packed);
  return cudaGetLastError() == cudaSuccess ? DRL_OK : DRL_ERR_CUDA;
}

// rms_in / rms_out device float64 [2d+4] (may not alias); obs_in/obs_out may alias; rew_in/rew_out may alias.
extern "C" int drl_vecnorm_apply(const float* obs_in, float* obs_out, const float* rew_in, float* rew_out, int32_t n,
                                 int32_t d, const double* packed, const double* rms_in, double* rms_out, float* ret,
                                 const uint8_t* done, float clip_obs, float clip_rew, float eps, int32_t flags,
                                 void* stream) {
  if (!obs_in || !obs_out || !rms_in || !rms_out || n <= 0 || d <= 0 || rms_in == rms_out) return DRL_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  long long work = (long long)n * d;
  int blocks = (int)((work + drl::kVnThreads * 4 - 1) / (drl::kVnThreads * 4));
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  drl::vecnorm_apply_kernel<<<blocks, drl::kVnThreads, (2 * d + 1) * sizeof(float), st>>>(
      obs_in, obs_out, rew_in, rew_out, n, d, packed, rms_in, rms_out, ret, done, clip_obs, clip_rew, eps, flags & 1,
      (flags >> 1) & 1, (flags >> 2) & 1);
  return cudaGetLastError() == cudaSuccess ? DRL_OK : DRL_ERR_CUDA;
}

// terminal observations of the environments that finished this step, normalised with the current statistics
// (what VecNormalize hands to SB3 in infos[i]["terminal_observation"]); rows of running environments are skipped
namespace drl {
__global__ void vecnorm_terminal_kernel(const float* __restrict__ tin, float* __restrict__ tout,
                                        const unsigned char* __restrict__ done, int n, int d,
                                        const double* __restrict__ rms, float clip_obs, float eps, int norm_obs) {
  const long long total = (long long)n * d;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e / d), col = (int)(e % d);
    if (!done[row]) continue;
    float x = tin[e];
    if (norm_obs)
      x = fminf(fmaxf((x - (float)rms[col]) * (float)(1.0 / sqrt(rms[d + col] + (double)eps)), -clip_obs), clip_obs);
    tout[e] = x;
  }
}
}  // namespace drl

extern "C" int drl_vecnorm_terminal(const float* tobs_in, float* tobs_out, const uint8_t* done, int32_t n, int32_t d,
                                    const double* rms, float clip_obs, float eps, int32_t norm_obs, void* stream) {
  if (!tobs_in || !tobs_out || !done || !rms || n <= 0 || d <= 0) return DRL_ERR_INVALID;
  long long work = (long long)n * d;
  int blocks = (int)((work + drl::kVnThreads * 4 - 1) / (drl::kVnThreads * 4));
  if (blocks > 148 * 8) blocks = 148 * 8;
  drl::vecnorm_terminal_kernel<<<blocks, drl::kVnThreads, 0, (cudaStream_t)stream>>>(tobs_in, tobs_out, done, n, d, rms,
                                                                                    clip_obs, eps, norm_obs);
  return cudaGetLastError() == cudaSuccess ? DRL_OK : DRL_ERR_CUDA;
}

// ---- FP32 roofline denominator: sustained FFMA rate of the device, measured (bench.py "fp32.peak") ----
namespace drl {
__global__ void __launch_bounds__(256) ffma_probe_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 4
  for (int i = 0; i < iters; i++) {     // 8 independent chains per thread: issue-bound, not latency-bound
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace drl

extern "C" int drl_fp32_peak_probe(int32_t device, double* tflops_out) {
  if (!tflops_out) return DRL_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return DRL_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DRL_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  float* buf = nullptr;
  if (cudaMalloc(&buf, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return DRL_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    drl::ffma_probe_kernel<<<blocks, threads>>>(buf, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return DRL_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0 && ms > 0.f) best = fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops_out = best;
  return DRL_OK;
}
