// Support points of the depth-2 homogeneous ICNN (HomogeneousICNN.forward, dair_pll/deep_support_function.py:238-266;
// called per contact by DeepSupportConvex.get_vertices, geometry.py:309-325) for ALL direction rows of a batch in one
// kernel on the sm_100a tensor cores: tcgen05.mma kind::i8 with int32 accumulators in TMEM, operands in shared memory,
// one thread per direction row for the prologue (slope-mask bits of layer 0) and the epilogue (fp64 reconstruction of
// the layer Jacobian, layer-1 mask, support point).  See cn_icnn_tc.cuh for the mathematics.
//
// Per CTA (128 threads, one per SM, persistent over 128-row tiles):
//   A operand   the tile's mask bits as bytes, twice: values {0, 1} and {0, -128}; K-major, no swizzle (64 KB)
//   B operand   digit planes of Q_k for a (chunk of 32 hidden units, k) unit: 6 planes x (32 x 256) int8 = 48 KB, copied
//               from the prepared image (L2-resident, 1.18 MB) by cp.async.bulk into a two-deep ring
//   TMEM        9 accumulators (3 k x 3 plane pairs) of 128 lanes x 32 columns
//   per chunk   3 units x 6 planes x 8 k-steps = 144 MMAs (M 128, N 32, K 32), one commit, then the epilogue of the
//               chunk's 32 hidden units straight out of TMEM (tcgen05.ld 32x32b)
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/dair_pll_b200.h"
#include "cn_icnn_tc.cuh"

namespace {

using namespace cn;

constexpr int kThreads = 128;
constexpr int kTileRows = 128;
constexpr int kSmemA = kTileRows * TC_W;                  // 32 KB per copy
constexpr int kOffAlo = 0, kOffAhi = kSmemA, kOffB = 2 * kSmemA, kOffC = kOffB + 2 * TC_UNIT_BYTES,
              kOffBar = kOffC + TC_NCONST * 8, kSmemBytes = kOffBar + 64;
constexpr uint32_t kTmemCols = 512;
// instruction descriptor, kind::i8: D s32 (2 << 4), A s8 (1 << 7), B s8 (1 << 10), both K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin (a protocol error must not hang the device): traps after ~2 s
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  // K-major, no swizzle: LBO (K-adjacent core matrices) 128 B, SBO (8-row groups) 2048 B, descriptor version 1
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// z1_i in plain fp64 for one row (the rare |z1| ~ 0 case, where a 42-bit Jacobian must not decide the mask)
__device__ __noinline__ double exact_z1(const double* sC, const double* __restrict__ Wh, double dx, double dy, double dz,
                                        double slope, int i) {
  const double* W0 = sC + TC_C_WD0;
  double z = 0;
  for (int j = 0; j < TC_W; ++j) {
    const double lin = dx * W0[j] + dy * W0[TC_W + j] + dz * W0[2 * TC_W + j];
    z += (lin > 0 ? lin : slope * lin) * fabs(Wh[j * TC_W + i]);
  }
  const double* W1 = sC + TC_C_WD1;
  return z + (dx * W1[i] + dy * W1[TC_W + i] + dz * W1[2 * TC_W + i]);
}

__global__ void __launch_bounds__(kThreads, 1)
icnn_tc_kernel(const double* __restrict__ d, int64_t D, const uint8_t* __restrict__ img, const double* __restrict__ consts,
               const double* __restrict__ Wh, double slope, double* __restrict__ p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  double* sC = reinterpret_cast<double*>(smem + kOffC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 48);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sAlo = smem_u32(smem + kOffAlo), sAhi = smem_u32(smem + kOffAhi), sB = smem_u32(smem + kOffB);
  const uint32_t bar_full0 = smem_u32(&bars[0]), bar_mma0 = smem_u32(&bars[2]), bar_chunk = smem_u32(&bars[4]);

  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < TC_NCONST; i += kThreads) sC[i] = consts[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);

  const int64_t ntiles = (D + kTileRows - 1) / kTileRows;
  uint32_t gu = 0;          // units issued so far by this CTA (identical in every thread)
  uint32_t chunks_done = 0;
  if (tid == 0 && (int64_t)blockIdx.x < ntiles) {
    mbar_expect_tx(bar_full0, TC_UNIT_BYTES);
    bulk_g2s(sB, img, TC_UNIT_BYTES, bar_full0);
  }
  const double* W0 = sC + TC_C_WD0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row = tile * kTileRows + tid;
    const bool valid = row < D;
    const bool last_tile = tile + gridDim.x >= ntiles;
    double dx = 0, dy = 0, dz = 0;
    if (valid) { dx = d[3 * row]; dy = d[3 * row + 1]; dz = d[3 * row + 2]; }
    // ---- prologue: slope-mask bits of layer 0 as the two byte-valued A operands ----
    {
      uint8_t* a_lo = smem + kOffAlo + (tid >> 3) * 2048 + (tid & 7) * 16;
      uint8_t* a_hi = smem + kOffAhi + (tid >> 3) * 2048 + (tid & 7) * 16;
#pragma unroll 1
      for (int jc = 0; jc < TC_W / 16; ++jc) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int j = jc * 16 + jj;
          const double lin = dx * W0[j] + dy * W0[TC_W + j] + dz * W0[2 * TC_W + j];
          w[jj >> 2] |= (lin > 0 ? 1u : 0u) << (8 * (jj & 3));
        }
        *reinterpret_cast<uint4*>(a_lo + jc * 128) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(a_hi + jc * 128) = make_uint4(w[0] << 7, w[1] << 7, w[2] << 7, w[3] << 7);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    double p0 = 0, p1 = 0, p2 = 0;
#pragma unroll 1
    for (int c = 0; c < TC_CHUNKS; ++c) {
      if (tid == 0) {
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
          const uint32_t b = gu & 1u;
          mbar_wait(bar_full0 + 8 * b, (gu >> 1) & 1u);
          tc_fence_after();
          const uint32_t sBu = sB + b * TC_UNIT_BYTES;
#pragma unroll 1
          for (int t = 0; t < TC_NACC; ++t) {
            const uint32_t acc = tmem + (uint32_t)((k * TC_NACC + t) * TC_NC);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t a_base = h == 0 ? sAhi : sAlo;
              const uint32_t b_base = sBu + (uint32_t)((2 * t + h) * TC_SLICE_BYTES);
#pragma unroll
              for (int ks = 0; ks < TC_W / 32; ++ks)
                umma_i8(acc, umma_desc(a_base + ks * 256), umma_desc(b_base + ks * 256), (h | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(bar_mma0 + 8 * b);
          if (k == 2) umma_commit(bar_chunk);
          // refill the other ring slot: its previous user (unit gu - 1) must have been read completely
          if (gu >= 1) mbar_wait(bar_mma0 + 8 * (b ^ 1u), ((gu - 1) >> 1) & 1u);
          const int next = (c * 3 + k + 1) % TC_UNITS;
          if (!(last_tile && c == TC_CHUNKS - 1 && k == 2)) {
            mbar_expect_tx(bar_full0 + 8 * (b ^ 1u), TC_UNIT_BYTES);
            bulk_g2s(sB + (b ^ 1u) * TC_UNIT_BYTES, img + (size_t)next * TC_UNIT_BYTES, TC_UNIT_BYTES, bar_full0 + 8 * (b ^ 1u));
          }
          ++gu;
        }
      } else {
        gu += 3;
      }
      __syncwarp();
      mbar_wait(bar_chunk, chunks_done & 1u);
      ++chunks_done;
      __syncwarp();
      tc_fence_after();
      // ---- epilogue: Jacobian Y_k, z1, layer-1 mask, support point -- 8 hidden units per TMEM read ----
#pragma unroll 1
      for (int g = 0; g < TC_NC / 8; ++g) {
        int32_t a[3 * TC_NACC][8];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 3 * TC_NACC; ++q) tmem_ld8(tmem_lane + (uint32_t)(q * TC_NC + g * 8), a[q]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = c * TC_NC + g * 8 + q;
          double Y[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            double v = (double)a[k * TC_NACC][q];
#pragma unroll
            for (int t = 1; t < TC_NACC; ++t) v = fma(v, 16384.0, (double)a[k * TC_NACC + t][q]);
            Y[k] = fma(sC[TC_C_COEF + k * TC_W + i], v, sC[TC_C_BASE + k * TC_W + i]);
          }
          double z = dx * Y[0] + dy * Y[1] + dz * Y[2];
          if (fabs(z) < sC[TC_C_ZTOL + i] && valid) z = exact_z1(sC, Wh, dx, dy, dz, slope, i);
          const double m = z > 0 ? sC[TC_C_WO + i] : slope * sC[TC_C_WO + i];
          p0 = fma(m, Y[0], p0);
          p1 = fma(m, Y[1], p1);
          p2 = fma(m, Y[2], p2);
        }
      }
      tc_fence_before();
      __syncthreads();          // TMEM accumulators (and, after the last chunk, the A operands) are free again
      tc_fence_after();
    }
    if (valid) { p[3 * row] = p0; p[3 * row + 1] = p1; p[3 * row + 2] = p2; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// one block per hidden unit i (column of the three Q_k): scales, digit planes, epilogue constants
__global__ void __launch_bounds__(TC_W)
icnn_tc_prepare_kernel(const double* __restrict__ Wd0, const double* __restrict__ Wd1, const double* __restrict__ Wh,
                       const double* __restrict__ wout, double slope, uint8_t* __restrict__ img, double* __restrict__ consts) {
  __shared__ double red_max[TC_W / 32], red_sum[TC_W / 32], red_abs[TC_W / 32];
  __shared__ double bc[3];
  const int i = blockIdx.x, j = threadIdx.x, lane = j & 31, w = j >> 5;
  const double wh = fabs(Wh[j * TC_W + i]);
  double abs_total = 0;
  for (int k = 0; k < 3; ++k) {
    const double q = Wd0[k * TC_W + j] * wh;
    double mx = fabs(q), sm = q, ab = fabs(q);
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
      ab += __shfl_xor_sync(0xffffffffu, ab, o);
    }
    if (lane == 0) { red_max[w] = mx; red_sum[w] = sm; red_abs[w] = ab; }
    __syncthreads();
    if (j == 0) {
      double m2 = 0, s2 = 0, a2 = 0;
      for (int x = 0; x < TC_W / 32; ++x) { m2 = fmax(m2, red_max[x]); s2 += red_sum[x]; a2 += red_abs[x]; }
      bc[0] = m2; bc[1] = s2; bc[2] = a2;
    }
    __syncthreads();
    int e;
    const double sigma = tc_column_scale(bc[0], &e);
    int8_t dig[TC_NS];
    tc_digits(q, e, dig);
    for (int s = 0; s < TC_NS; ++s) img[tc_image_offset(k, s, j, i)] = (uint8_t)dig[s];
    if (j == 0) {
      consts[TC_C_BASE + k * TC_W + i] = Wd1[k * TC_W + i] + slope * bc[1];
      consts[TC_C_COEF + k * TC_W + i] = (1.0 - slope) * ldexp(sigma, -(7 * TC_NS - 1));
      consts[TC_C_WD0 + k * TC_W + i] = Wd0[k * TC_W + i];
      consts[TC_C_WD1 + k * TC_W + i] = Wd1[k * TC_W + i];
    }
    abs_total += bc[2] + fabs(Wd1[k * TC_W + i]);
    __syncthreads();
  }
  if (j == 0) {
    consts[TC_C_WO + i] = fabs(wout[i]);
    consts[TC_C_ZTOL + i] = TC_ZTOL_REL * abs_total;
  }
}

}  // namespace

extern "C" {

size_t dpll_icnn_tc_image_bytes(void) { return (size_t)cn::TC_IMG_BYTES; }
size_t dpll_icnn_tc_const_bytes(void) { return (size_t)cn::TC_NCONST * sizeof(double); }

int dpll_icnn_tc_prepare_f64(const double* Wd0, const double* Wd1, const double* Wh, const double* wout, int32_t W,
                             double slope, void* image, double* consts, void* stream) {
  if (W != cn::TC_W || !Wd0 || !Wd1 || !Wh || !wout || !image || !consts) return DPLL_EINVAL;
  icnn_tc_prepare_kernel<<<cn::TC_W, cn::TC_W, 0, static_cast<cudaStream_t>(stream)>>>(
      Wd0, Wd1, Wh, wout, slope, static_cast<uint8_t*>(image), consts);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPLL_OK : (int)e;
}

int dpll_icnn_tc_support_f64(const double* d, int64_t D, const void* image, const double* consts, const double* Wh,
                             int32_t W, double slope, double* p, void* stream) {
  if (W != cn::TC_W || D < 0 || !image || !consts || !Wh) return DPLL_EINVAL;
  if (D == 0) return DPLL_OK;
  if (!d || !p) return DPLL_EINVAL;
  cudaError_t e = cudaFuncSetAttribute(icnn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (D + kTileRows - 1) / kTileRows;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  icnn_tc_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      d, D, static_cast<const uint8_t*>(image), consts, Wh, slope, p);
  e = cudaGetLastError();
  return e == cudaSuccess ? DPLL_OK : (int)e;
}

}  // extern "C"
