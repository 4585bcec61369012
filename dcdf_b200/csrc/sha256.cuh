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
// digest[j] = SHA2-256(prefix_j ++ blob[off_j .. off_j + len_j))
__global__ void k_sha256(const u8* blob, const HashJob* jobs, u32 n_jobs, u8* digests) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  const HashJob job = jobs[j];
  Sha256 s;
  s.init();
  for (u32 i = 0; i < job.n_prefix; i++) s.byte(job.prefix[i]);
  const u8* p = blob + job.off;
  u64 i = 0;
  // head: bytes until the hash is at a block boundary and nothing else matters; then whole blocks from unaligned
  // source bytes (two aligned 32-bit loads + a funnel shift per word, as the decoders read big-endian fields)
  while (i < job.len && (s.n & 63) != 0) s.byte(p[i++]);
  const uintptr_t a0 = (uintptr_t)(p + i);
  const u32 sh = (u32)(a0 & 3) * 8u;
  const u32* wp = reinterpret_cast<const u32*>(a0 & ~(uintptr_t)3);
  while (i + 64 <= job.len) {
    u32 be[16];
    u32 lo = wp[0];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const u32 hi = sh ? wp[k + 1] : 0u;  // blobs carry 64 bytes of padding, reading one word past the body is fine
      const u32 le = sh ? __funnelshift_r(lo, hi, sh) : lo;
      be[k] = __byte_perm(le, 0, 0x0123);
      lo = sh ? hi : wp[k + 1];
    }
    s.block(be);
    wp += 16;
    i += 64;
  }
  while (i < job.len) s.byte(p[i++]);
  s.finish(digests + 32ull * j);
}
#endif

}  // namespace dcdf
