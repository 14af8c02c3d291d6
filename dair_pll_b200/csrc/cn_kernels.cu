// CUDA kernels + C ABI (include/dair_pll_b200.h) for the ContactNets hot path, sm_100a.
//
// One sample per thread: the whole per-sample pipeline (terms -> QP build -> Newton solve
// -> loss -> envelope backward) lives in registers; no tensor cores (the per-sample
// systems are 6x6), no shared-memory tiles; HBM traffic is 2*13 loads + 1 store per
// sample.  Parameter gradients are accumulated per thread across a grid-stride loop and
// reduced warp -> block -> grid in a fixed order (deterministic).
#include <cuda_runtime.h>

#include "../../include/dair_pll_b200.h"
#include "cn_cube.cuh"

namespace {

constexpr int kLossThreads = 128;
constexpr int kNAcc = 16;            // 14 parameter gradients + loss sum + pad
constexpr int kMaxBlocks = 148 * 16; // upper bound used to size the workspace

struct DeviceInfo { int sms; int device; };
inline DeviceInfo device_info() {
  int dev = 0;
  cudaGetDevice(&dev);
  static int cached_dev = -1, cached_sms = 0;
  if (cached_dev != dev) {
    cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return {cached_sms, dev};
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads)
cube_loss_kernel(const T* __restrict__ x, const T* __restrict__ xp, const T* __restrict__ weight,
                 const T* __restrict__ inertia,
                 const T* __restrict__ mu, const T* __restrict__ half, T dt, T eps, int64_t B,
                 T* __restrict__ loss, T* __restrict__ force, int32_t* __restrict__ iters,
                 T* __restrict__ partials, int want_grad) {
  cn::CubeParams<T> P;
  cn::cube_params_init(P, inertia, mu, half, dt, eps);
  const cn::SolverCfg<T> cfg = cn::default_cfg<T>();
  T acc[kNAcc];
#pragma unroll
  for (int i = 0; i < kNAcc; ++i) acc[i] = T(0);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
    T xs[13], xps[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) { xs[i] = x[b * 13 + i]; xps[i] = xp[b * 13 + i]; }
    int it;
    T gs[DPLL_CUBE_NPARAM];
#pragma unroll
    for (int i = 0; i < DPLL_CUBE_NPARAM; ++i) gs[i] = T(0);
    const T l = cn::cube_loss_sample<T>(P, cfg, xs, xps, want_grad ? gs : nullptr,
                                        force ? force + b * 12 : nullptr, &it);
    const T w = weight ? weight[b] : T(1);
#pragma unroll
    for (int i = 0; i < DPLL_CUBE_NPARAM; ++i) acc[i] += w * gs[i];
    loss[b] = l;
    acc[14] += l;
    if (iters) iters[b] = it;
  }
  if (!partials) return;
  __shared__ T red[kLossThreads / 32][kNAcc];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kNAcc; ++i) {
    const T s = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < kNAcc) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) s += red[w][threadIdx.x];
    partials[(int64_t)blockIdx.x * kNAcc + threadIdx.x] = s;
  }
}

// Fixed-order reduction of the per-block partials: warp w owns accumulator w.
template <typename T>
__global__ void reduce_partials_kernel(const T* __restrict__ partials, int nblocks, T* __restrict__ grad,
                                       T* __restrict__ loss_sum) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w >= kNAcc) return;
  T s = T(0);
  for (int b = lane; b < nblocks; b += 32) s += partials[(int64_t)b * kNAcc + w];
  s = warp_sum(s);
  if (lane == 0) {
    if (w < DPLL_CUBE_NPARAM) { if (grad) grad[w] = s; }
    else if (w == 14) { if (loss_sum) *loss_sum = s; }
  }
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads)
cube_rollout_kernel(const T* __restrict__ x0, const T* __restrict__ inertia, const T* __restrict__ mu,
                    const T* __restrict__ half, T dt, T eps, int64_t B, int steps, T* __restrict__ traj,
                    T* __restrict__ force, int32_t* __restrict__ iters) {
  cn::CubeParams<T> P;
  cn::cube_params_init(P, inertia, mu, half, dt, eps);
  const cn::SolverCfg<T> cfg = cn::default_cfg<T>();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
    T xc[13], xn[13];
    T* out = traj + b * (int64_t)(steps + 1) * 13;
#pragma unroll
    for (int i = 0; i < 13; ++i) { xc[i] = x0[b * 13 + i]; out[i] = xc[i]; }
    int total = 0;
    for (int s = 0; s < steps; ++s) {
      total += cn::cube_step_sample<T>(P, cfg, xc, xn, force ? force + (b * steps + s) * 12 : nullptr);
#pragma unroll
      for (int i = 0; i < 13; ++i) { xc[i] = xn[i]; out[(int64_t)(s + 1) * 13 + i] = xn[i]; }
    }
    if (iters) iters[b] = total;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int64_t iters) {
  T a[16];
  const T m = T(1) + T(1e-9) * T(threadIdx.x), c = T(1e-12);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = T(i) + T(blockIdx.x) * T(1e-6);
  for (int64_t k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = a[i] * m + c;
  }
  T s = T(0);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T>
int launch_cube_loss(const T* x, const T* xp, const T* weight, const T* inertia, const T* mu, const T* half, T dt, T eps,
                     int64_t B, T* loss, T* force, int32_t* iters, T* grad, T* loss_sum, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (B < 0 || !inertia || !mu || !half) return DPLL_EINVAL;
  if (B > 0 && (!x || !xp || !loss)) return DPLL_EINVAL;
  const bool want_red = grad || loss_sum;
  if (want_red && (!workspace || workspace_bytes < dpll_workspace_bytes())) return DPLL_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DeviceInfo di = device_info();
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cube_loss_kernel<T>, kLossThreads, 0);
  if (per_sm < 1) per_sm = 1;
  int64_t need = (B + kLossThreads - 1) / kLossThreads;
  int64_t cap = (int64_t)di.sms * per_sm;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  int blocks = (int)(need < cap ? need : cap);
  if (blocks < 1) blocks = 1;
  T* partials = want_red ? static_cast<T*>(workspace) : nullptr;
  cube_loss_kernel<T><<<blocks, kLossThreads, 0, st>>>(x, xp, weight, inertia, mu, half, dt, eps, B, loss, force,
                                                       iters, partials, grad ? 1 : 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (want_red) {
    reduce_partials_kernel<T><<<1, 32 * kNAcc, 0, st>>>(partials, blocks, grad, loss_sum);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return DPLL_OK;
}

template <typename T>
int launch_cube_rollout(const T* x0, const T* inertia, const T* mu, const T* half, T dt, T eps, int64_t B,
                        int32_t steps, T* traj, T* force, int32_t* iters, void* stream) {
  if (B < 0 || steps < 0 || !inertia || !mu || !half) return DPLL_EINVAL;
  if (B > 0 && (!x0 || !traj)) return DPLL_EINVAL;
  if (B == 0) return DPLL_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DeviceInfo di = device_info();
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cube_rollout_kernel<T>, kLossThreads, 0);
  if (per_sm < 1) per_sm = 1;
  int64_t need = (B + kLossThreads - 1) / kLossThreads;
  int64_t cap = (int64_t)di.sms * per_sm;
  int blocks = (int)(need < cap ? need : cap);
  cube_rollout_kernel<T><<<blocks, kLossThreads, 0, st>>>(x0, inertia, mu, half, dt, eps, B, steps, traj, force,
                                                          iters);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPLL_OK : (int)e;
}

}  // namespace

extern "C" {

int dpll_version(void) { return 100; }

size_t dpll_workspace_bytes(void) { return (size_t)kMaxBlocks * kNAcc * sizeof(double); }

int dpll_cube_loss_f64(const double* x, const double* x_plus, const double* weight, const double* inertia,
                       const double* mu_pair, const double* half, double dt, double eps, int64_t B, double* loss, double* force,
                       int32_t* iters, double* grad, double* loss_sum, void* workspace, size_t workspace_bytes,
                       void* stream) {
  return launch_cube_loss<double>(x, x_plus, weight, inertia, mu_pair, half, dt, eps, B, loss, force, iters, grad,
                                  loss_sum, workspace, workspace_bytes, stream);
}

int dpll_cube_loss_f32(const float* x, const float* x_plus, const float* weight, const float* inertia,
                       const float* mu_pair, const float* half, float dt, float eps, int64_t B, float* loss, float* force,
                       int32_t* iters, float* grad, float* loss_sum, void* workspace, size_t workspace_bytes,
                       void* stream) {
  return launch_cube_loss<float>(x, x_plus, weight, inertia, mu_pair, half, dt, eps, B, loss, force, iters, grad,
                                 loss_sum, workspace, workspace_bytes, stream);
}

int dpll_cube_rollout_f64(const double* x0, const double* inertia, const double* mu_pair, const double* half,
                          double dt, double eps, int64_t B, int32_t steps, double* traj, double* force,
                          int32_t* iters, void* stream) {
  return launch_cube_rollout<double>(x0, inertia, mu_pair, half, dt, eps, B, steps, traj, force, iters, stream);
}

int dpll_cube_rollout_f32(const float* x0, const float* inertia, const float* mu_pair, const float* half, float dt,
                          float eps, int64_t B, int32_t steps, float* traj, float* force, int32_t* iters,
                          void* stream) {
  return launch_cube_rollout<float>(x0, inertia, mu_pair, half, dt, eps, B, steps, traj, force, iters, stream);
}

int dpll_fma_peak_f64(double* out, int32_t blocks, int64_t iters, void* stream) {
  if (!out || blocks <= 0 || iters < 0) return DPLL_EINVAL;
  fma_peak_kernel<double><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, iters);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPLL_OK : (int)e;
}

int dpll_fma_peak_f32(float* out, int32_t blocks, int64_t iters, void* stream) {
  if (!out || blocks <= 0 || iters < 0) return DPLL_EINVAL;
  fma_peak_kernel<float><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, iters);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPLL_OK : (int)e;
}

}  // extern "C"
