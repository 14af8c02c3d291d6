// Support-function network forward on the 5th-generation tensor cores (tcgen05 / TMEM, sm_100a): constants and the
// host/device weight-slicing arithmetic of csrc/cn_icnn_tc.cu.
//
// HomogeneousICNN (dair_pll/deep_support_function.py:238-266), depth 2, width W = 256, LeakyReLU slope s:
//   h0 = lrelu(d Wd0),  z1 = h0 |Wh| + d Wd1,  f = |wout| . lrelu(z1),  support point p = d f / d d.
// z1 is piecewise LINEAR in d, so with the slope-mask bits b0[j] = (d.Wd0[:,j] > 0):
//   d z1_i / d d_k = Y_k[i] = Wd1[k,i] + s * sum_j Q_k[j,i] + (1 - s) * sum_j b0[j] Q_k[j,i],   Q_k[j,i] = Wd0[k,j] |Wh|[j,i]
//   z1_i = sum_k d_k Y_k[i]   (homogeneity),     p_k = sum_i |wout|_i lrelu'(z1_i) Y_k[i]
// -- the reference's reverse-mode recursion (:251-264) evaluated in forward mode.  The only batch-sized contraction
// left is  (b0 : D x 256, BINARY) x (Q_k : 256 x 256, constant), three times.  A binary operand is exact in int8, and
// a constant fp64 operand splits exactly into TC_NS balanced base-128 digits per column scale,
//   Q_k[j,i] ~= sigma_ki 2^-(7 NS - 1) sum_s dig_s[j,i] 128^(NS-1-s),   dig_s in [-64, 64]   (error <= 2^-(7 NS) sigma_ki),
// so each digit plane is an exact int8 x int8 -> int32 tensor-core product.  Two planes share one accumulator: the
// A operand exists as b0 (values 0 / 1) and as -128 b0 (values 0 / -128), the even planes are stored negated, and
//   acc_t = sum_j b0[j] (128 dig_{2t}[j,i] + dig_{2t+1}[j,i])        (|acc_t| < 2^23).
// The epilogue (one thread per direction row = one TMEM lane) rebuilds Y_k in fp64 by Horner's rule (exact integers
// below 2^53) and finishes z1, the mask and p in registers: nothing of size D x 256 ever reaches HBM.
#pragma once
#include <cstdint>
#include <cmath>

#ifdef __CUDACC__
#define TC_HD __host__ __device__ inline
#else
#define TC_HD inline
#endif

namespace cn {

constexpr int TC_W = 256;                            // network width handled by the tensor-core kernel
constexpr int TC_NS = 6;                             // digit planes per weight (7 bits each: 42-bit fixed point per column)
constexpr int TC_NACC = TC_NS / 2;                   // accumulators per input coordinate k
constexpr int TC_NC = 32;                            // output columns (hidden units i) per unit = N of one MMA
constexpr int TC_CHUNKS = TC_W / TC_NC;              // 8
constexpr int TC_UNITS = 3 * TC_CHUNKS;              // (chunk, k) units per 128-row tile: 24
constexpr int TC_SLICE_BYTES = TC_NC * TC_W;         // one digit plane of one unit: N x K int8 = 8 KB
constexpr int TC_UNIT_BYTES = TC_NS * TC_SLICE_BYTES;          // 48 KB
constexpr int TC_IMG_BYTES = TC_UNITS * TC_UNIT_BYTES;         // 1,179,648 B: all planes, in shared-memory image order
// constants (doubles): per hidden unit i eight values [coef_0, base_0, coef_1, base_1, coef_2, base_2, ztol, |wout|_i] with the
// output weight |wout|_i folded in (Y'_k = |wout|_i Y_k = base_k + coef_k * integer), then Wd0[3][W], Wd1[3][W]
constexpr int TC_C_COL = 0, TC_C_WD0 = 8 * TC_W, TC_C_WD1 = 11 * TC_W, TC_NCONST = 14 * TC_W;
constexpr int TC_NCONST_SMEM = 11 * TC_W;            // what the kernel keeps in shared memory (Wd1 is only read by the rare-path correction, redo_entries)
constexpr double TC_ZTOL_REL = 1e-9;                 // |z1| below this fraction of its scale is re-evaluated in fp64

// K-major, no-swizzle UMMA canonical layout of an (rows x 256 B) int8 operand: 8-row x 16-byte core matrices,
// K-adjacent core matrices 128 B apart (LBO), 8-row groups 2048 B apart (SBO)
TC_HD int tc_operand_offset(int row, int kbyte) {
  return (row >> 3) * 2048 + (kbyte >> 4) * 128 + (row & 7) * 16 + (kbyte & 15);
}

// byte offset of digit plane s of Q_k[j, i] in the image
TC_HD int tc_image_offset(int k, int s, int j, int i) {
  const int unit = (i / TC_NC) * 3 + k;
  return unit * TC_UNIT_BYTES + s * TC_SLICE_BYTES + tc_operand_offset(i % TC_NC, j);
}

// power-of-two column scale: cmax < sigma <= 2 cmax  (sigma = 1 for an all-zero column)
TC_HD double tc_column_scale(double cmax, int* e_out) {
  int e = 0;
  if (cmax > 0) (void)frexp(cmax, &e);
  *e_out = e;
  return ldexp(1.0, e);
}

// balanced base-128 digits of q / sigma (|q| < sigma = 2^e), most significant first; dig[0] in [-64, 64], the rest in
// [-64, 63].  Stored form: even planes negated (they multiply the -128 copy of the binary operand).
TC_HD void tc_digits(double q, int e, int8_t* dig) {
  long long I = llrint(ldexp(q, 7 * TC_NS - 1 - e));
  for (int s = TC_NS - 1; s >= 1; --s) {
    const int dd = (int)((I + 64) & 127) - 64;
    dig[s] = (int8_t)dd;
    I = (I - dd) >> 7;
  }
  dig[0] = (int8_t)I;
  for (int s = 0; s < TC_NS; s += 2) dig[s] = (int8_t)(-dig[s]);
}

// the value the tensor-core path works with (for tests): sigma 2^-(7 NS - 1) sum_s dig_s 128^(NS-1-s)
TC_HD double tc_reconstruct(const int8_t* dig, int e) {
  double v = 0;
  for (int s = 0; s < TC_NS; ++s) v = v * 128.0 + (double)((s & 1) ? dig[s] : -dig[s]);
  return ldexp(v, e - (7 * TC_NS - 1));
}

}  // namespace cn
