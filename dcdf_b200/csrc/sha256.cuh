// sha256.cuh -- SHA2-256 (FIPS 180-4) for content addressing of stored nodes.
//
// The reference's stores hash every saved node; its test mapper builds a CIDv1 from a SHA2-256 multihash
// (testing.rs:172-183).  Chunk nodes are the bulk of the bytes and already sit in HBM, so they are hashed there: one
// thread per chunk (a SHA-256 message is a sequential chain of 64-byte blocks; the parallelism is across the tens of
// thousands of chunks of a build).  The same compression function, compiled for the host, hashes the small Links and
// superchunk nodes.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace dcdf {

struct Sha256 {
  uint32_t h[8];
  uint32_t w[16];   // current block, big-endian words
  uint64_t n;       // bytes so far

  __host__ __device__ static constexpr uint32_t rotr(uint32_t x, int s) { return (x >> s) | (x << (32 - s)); }
  __host__ __device__ static uint32_t k(int i) {
    constexpr uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    return K[i];
  }
  __host__ __device__ void init() {
    h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
    h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
    n = 0;
    for (int i = 0; i < 16; i++) w[i] = 0;
  }
  __host__ __device__ void compress() {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = w[i];
#pragma unroll 16
    for (int i = 0; i < 64; i++) {
      if (i >= 16) {
        const uint32_t w15 = m[(i + 1) & 15], w2 = m[(i + 14) & 15];
        const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3), s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
        m[i & 15] = m[i & 15] + s0 + m[(i + 9) & 15] + s1;
      }
      const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
      const uint32_t t1 = hh + S1 + ch + k(i) + m[i & 15];
      const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
      const uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  __host__ __device__ void byte(uint8_t v) {
    const int pos = (int)(n & 63);
    const int wi = pos >> 2, sh = 24 - 8 * (pos & 3);
    w[wi] = (w[wi] & ~(0xffu << sh)) | ((uint32_t)v << sh);
    n++;
    if ((n & 63) == 0) compress();
  }
  __host__ __device__ void update(const uint8_t* p, uint64_t len) {
    for (uint64_t i = 0; i < len; i++) byte(p[i]);
  }
  // 16 big-endian words at once (n must be a multiple of 64)
  __host__ __device__ void block(const uint32_t (&be)[16]) {
    for (int i = 0; i < 16; i++) w[i] = be[i];
    n += 64;
    compress();
  }
  __host__ __device__ void finish(uint8_t out[32]) {
    const uint64_t bits = n * 8;
    byte(0x80);
    while ((n & 63) != 56) byte(0);
    for (int i = 7; i >= 0; i--) byte((uint8_t)(bits >> (8 * i)));
    for (int i = 0; i < 8; i++) {
      out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
      out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
    }
  }
};

#ifdef __CUDACC__
struct HashJob {
  u64 off;        // message body inside `blob`
  u64 len;
  u8 prefix[8];   // bytes hashed before the body (node header, resolver.rs:130-132 + mmstruct.rs:215-218)
  u32 n_prefix;
  u32 pad_;
};

// One compression (FIPS 180-4 6.2.2), fully unrolled: the round constants become immediates, the message schedule
// stays in sixteen registers.
DCDF_DEVINL void sha256_compress_dev(u32 (&h)[8], u32 (&m)[16]) {
  constexpr u32 K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
      0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
      0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
      0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
      0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
      0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    if (i >= 16) {
      const u32 w15 = m[(i + 1) & 15], w2 = m[(i + 14) & 15];
      const u32 s0 = __funnelshift_r(w15, w15, 7) ^ __funnelshift_r(w15, w15, 18) ^ (w15 >> 3);
      const u32 s1 = __funnelshift_r(w2, w2, 17) ^ __funnelshift_r(w2, w2, 19) ^ (w2 >> 10);
      m[i & 15] = m[i & 15] + s0 + m[(i + 9) & 15] + s1;
    }
    const u32 S1 = __funnelshift_r(e, e, 6) ^ __funnelshift_r(e, e, 11) ^ __funnelshift_r(e, e, 25), ch = (e & f) ^ (~e & g);
    const u32 t1 = hh + S1 + ch + K[i] + m[i & 15];
    const u32 S0 = __funnelshift_r(a, a, 2) ^ __funnelshift_r(a, a, 13) ^ __funnelshift_r(a, a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
    const u32 t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// digest[j] = SHA2-256(prefix_j ++ blob[off_j .. off_j + len_j)).  One thread per message (a SHA-256 message is a
// sequential chain of 64-byte blocks); the message is walked block by block with a single call site of the
// compression: blocks that lie inside the body are read as aligned word pairs + funnel shift, the first and the last
// ones (prefix, 0x80, zero fill, bit length) byte by byte.
__global__ void __launch_bounds__(32) k_sha256(const u8* blob, const HashJob* jobs, u32 n_jobs, u8* digests) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const HashJob job = jobs[j];
  const u8* body = blob + job.off;
  const u64 npre = job.n_prefix, total = npre + job.len;
  const u64 n_blocks = (total + 9ull + 63ull) / 64ull;
  u32 h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
#pragma unroll 1
  for (u64 blk = 0; blk < n_blocks; blk++) {
    const u64 pos = 64ull * blk;  // message offset of the block
    u32 m[16];
    if (pos >= npre && pos + 64ull <= total) {
      const uintptr_t a0 = (uintptr_t)(body + (pos - npre));
      const u32 sh = (u32)(a0 & 3) * 8u;
      const u32* wp = reinterpret_cast<const u32*>(a0 & ~(uintptr_t)3);
      u32 lo = wp[0];
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const u32 hi = sh ? wp[k + 1] : 0u;  // blobs carry 64 bytes of padding, reading one word past the body is fine
        const u32 le = sh ? __funnelshift_r(lo, hi, sh) : lo;
        m[k] = __byte_perm(le, 0, 0x0123);
        lo = sh ? hi : wp[k + 1];
      }
    } else {
#pragma unroll 1
      for (int k = 0; k < 16; k++) {
        u32 w = 0;
#pragma unroll 1
        for (int t = 0; t < 4; t++) {
          const u64 q = pos + (u64)(4 * k + t);
          u32 v = 0;
          if (q < npre) v = job.prefix[q];
          else if (q < total) v = body[q - npre];
          else if (q == total) v = 0x80u;
          else if (q >= 64ull * n_blocks - 8ull) v = (u32)(((total * 8ull) >> (8ull * (64ull * n_blocks - 1ull - q))) & 0xffull);
          w = (w << 8) | v;
        }
        // m[k] = w without a dynamically indexed register array
#pragma unroll
        for (int z = 0; z < 16; z++) if (z == k) m[z] = w;
      }
    }
    sha256_compress_dev(h, m);
  }
  u8* out = digests + 32ull * j;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (u8)(h[i] >> 24); out[4 * i + 1] = (u8)(h[i] >> 16);
    out[4 * i + 2] = (u8)(h[i] >> 8); out[4 * i + 3] = (u8)h[i];
  }
}
#endif

}  // namespace dcdf
