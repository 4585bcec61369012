// common.cuh -- device helpers shared by the encode / decode kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <limits>

namespace dcdf {

typedef int64_t i64;
typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

// Error bits accumulated in a device flag word, checked once per call on the host.
enum : u32 {
  EF_NONFINITE = 1u << 0,
  EF_PRECISION = 1u << 1,
  EF_OVERFLOW = 1u << 2,
  EF_ARENA_FULL = 1u << 3,
  EF_BAD_FORMAT = 1u << 4,
  EF_OUT_CAP = 1u << 5,
  EF_BAD_LEVELS = 1u << 6,    // superchunk.rs:105-110 inside a recursion
  EF_REGION_EXACT = 1u << 7,  // nested region needs the exact saturating-cast pass (not built yet)
  EF_OUT_OF_BOUNDS = 1u << 8, // a device-resident query outside the array (mmarray.rs:218-229)
};

#define DCDF_DEVINL __device__ __forceinline__

// ------------------------------------------------------------------ fixed point (fixed.rs:31-86)
// to_fixed runs in the INPUT float type (mmbuffer.rs:565-571 passes the f32 straight in).
template <typename F>
DCDF_DEVINL i64 to_fixed_dev(F n, int bits, bool round, u32& err) {
  if (n != n) return 0;  // NaN -> 0  (fixed.rs:35-37)
  if (isinf(n)) {  // +-inf  (fixed.rs:39-41)
    err |= EF_NONFINITE;
    return 0;
  }
  F shifted = n * (F)(i64(1) << bits);  // exact power-of-two scaling (may overflow to inf)
  F tr = sizeof(F) == 4 ? (F)truncf((float)shifted) : (F)trunc((double)shifted);
  if (shifted - tr > F(0)) {  // positive fractions only (fixed.rs:47)
    if (round) {
      shifted = sizeof(F) == 4 ? (F)roundf((float)shifted) : (F)::round((double)shifted);  // half away from zero
    } else {
      err |= EF_PRECISION;
    }
  }
  shifted = shifted * F(2);
  const F lim = (F)9223372036854775808.0;
  if (!(shifted >= -lim && shifted < lim)) {  // num-traits to_i64() == None  (fixed.rs:66-69)
    err |= EF_OVERFLOW;
    return 0;
  }
  i64 v = sizeof(F) == 4 ? __float2ll_rz((float)shifted) : __double2ll_rz((double)shifted);
  return v + 1;
}

template <typename F>
DCDF_DEVINL F from_fixed_dev(i64 n, int bits) {
  if (n == 0) return sizeof(F) == 4 ? (F)__int_as_float(0x7fc00000) : (F)__longlong_as_double(0x7ff8000000000000ll);
  // i64 -> F is round-to-nearest-even, then an exact power-of-two divide (fixed.rs:84)
  F num = sizeof(F) == 4 ? (F)__ll2float_rn(n - 1) : (F)__ll2double_rn(n - 1);
  return num / (F)(i64(1) << (bits + 1));
}

// Input element -> i64 as MMBuffer3::get does (mmbuffer.rs:303-310).
template <typename T>
struct Conv;
template <>
struct Conv<float> {
  static DCDF_DEVINL i64 get(float v, int bits, bool round, u32& err) { return to_fixed_dev<float>(v, bits, round, err); }
};
template <>
struct Conv<double> {
  static DCDF_DEVINL i64 get(double v, int bits, bool round, u32& err) { return to_fixed_dev<double>(v, bits, round, err); }
};
template <>
struct Conv<int32_t> {
  static DCDF_DEVINL i64 get(int32_t v, int, bool, u32&) { return (i64)v; }
};
template <>
struct Conv<i64> {
  static DCDF_DEVINL i64 get(i64 v, int, bool, u32&) { return v; }
};

// ------------------------------------------------------------------ zigzag + byte length (dac.rs:134-142)
DCDF_DEVINL u64 zigzag64(i64 n) { return (u64)((n >> 63) ^ (i64)((u64)n << 1)); }
DCDF_DEVINL i64 unzigzag64(u64 zz) { return (i64)((zz >> 1) ^ (0 - (zz & 1))); }
// Valid when |n| < 2^31 (the narrow path guarantees |n| < 2^30).
DCDF_DEVINL u32 zigzag32(int32_t n) { return (u32)((n >> 31) ^ (int32_t)((u32)n << 1)); }
// Number of DAC bytes of a zigzag code: 0 still takes one byte (dac.rs:109-121).
DCDF_DEVINL int dac_len(u64 zz) { return zz == 0 ? 1 : (71 - __clzll((long long)zz)) >> 3; }
DCDF_DEVINL int dac_len(u32 zz) { return zz == 0 ? 1 : (39 - __clz((int)zz)) >> 3; }

// ------------------------------------------------------------------ sizes (bitmap.rs:166-172, dac.rs:66-75)
__host__ __device__ inline u32 bitmap_size(u32 length) { return 8u + 4u * (length / 128u) + 4u * ((length + 31u) / 32u); }

// ------------------------------------------------------------------ big-endian stores (extio.rs:196-249)
DCDF_DEVINL void store_be32(u8* p, u32 v) {
  p[0] = (u8)(v >> 24);
  p[1] = (u8)(v >> 16);
  p[2] = (u8)(v >> 8);
  p[3] = (u8)v;
}
DCDF_DEVINL u32 load_be32(const u8* p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3]; }

// ------------------------------------------------------------------ Morton helpers (row bit above column bit)
// Compact the even bits of a 16-bit Morton code: x0 y0 x1 y1 ... -> x
DCDF_DEVINL u32 compact1by1(u32 v) {
  v &= 0x55555555u;
  v = (v ^ (v >> 1)) & 0x33333333u;
  v = (v ^ (v >> 2)) & 0x0f0f0f0fu;
  v = (v ^ (v >> 4)) & 0x00ff00ffu;
  v = (v ^ (v >> 8)) & 0x0000ffffu;
  return v;
}
DCDF_DEVINL u32 morton_col(u32 m) { return compact1by1(m); }
DCDF_DEVINL u32 morton_row(u32 m) { return compact1by1(m >> 1); }
DCDF_DEVINL u32 spread1by1(u32 v) {
  v &= 0x0000ffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
DCDF_DEVINL u32 morton_encode(u32 row, u32 col) { return (spread1by1(row) << 1) | spread1by1(col); }

DCDF_DEVINL u32 lanemask_lt() {
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace dcdf
