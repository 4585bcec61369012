// xfer.cuh -- small host <-> device transfers as kernels over mapped pinned memory.
//
// The copy engines serve one bulk transfer after the other: a 100 KB table upload queued behind another context's
// 2 GB raster upload waits for all of it, and with it every kernel of the call.  Tables, flags and counters
// therefore move through pinned host memory that the GPU reads / writes directly (unified addressing maps every
// cudaMallocHost allocation into the device address space); only rasters and encoded blobs use cudaMemcpyAsync.
#pragma once
#include <algorithm>

#include "host.hpp"

namespace dcdf {

static __global__ void k_copy_bytes(u8* __restrict__ dst, const u8* __restrict__ src, size_t n) {
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
  if (((((uintptr_t)dst) | ((uintptr_t)src)) & 15u) == 0) {
    const size_t n16 = n >> 4;
    for (size_t i = tid; i < n16; i += nt) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = (n16 << 4) + tid; i < n; i += nt) dst[i] = src[i];
  } else {
    for (size_t i = tid; i < n; i += nt) dst[i] = src[i];
  }
}

// dst / src: device memory or pinned host memory, any alignment.
inline void copy_by_kernel(dcdf_ctx* ctx, void* dst, const void* src, size_t n) {
  if (n == 0) return;
  const size_t blocks = std::min<size_t>(std::max<size_t>(1, (n / 16 + 255) / 256), (size_t)4 * ctx->sm_count);
  k_copy_bytes<<<(unsigned)blocks, 256, 0, ctx->stream>>>(static_cast<u8*>(dst), static_cast<const u8*>(src), n);
  CK(cudaGetLastError());
}

// Host (pageable or pinned) -> device, ordered on ctx->stream; `src` may be reused as soon as the call returns.
inline void upload_small(dcdf_ctx* ctx, void* dev_dst, const void* src, size_t n) {
  if (n == 0) return;
  const size_t need = (n + 15) & ~size_t(15);
  if (ctx->up_off + need > ctx->xfer_up.cap) {
    CK(cudaStreamSynchronize(ctx->stream));  // nothing may still be reading the staging area
    ctx->up_off = 0;
    ctx->xfer_up.reserve(std::max<size_t>(need, 1u << 20));
  }
  u8* p = ctx->xfer_up.as<u8>() + ctx->up_off;
  memcpy(p, src, n);
  ctx->up_off += need;
  copy_by_kernel(ctx, dev_dst, p, n);
}

void sync_reads(dcdf_ctx* ctx);

// Device -> host, completed by the next sync_reads(ctx).
inline void read_small(dcdf_ctx* ctx, void* host_dst, const void* dev_src, size_t n) {
  if (n == 0) return;
  const size_t need = (n + 15) & ~size_t(15);
  if (ctx->down_off + need > ctx->xfer_down.cap) {
    sync_reads(ctx);
    ctx->xfer_down.reserve(std::max<size_t>(need, 1u << 20));
  }
  copy_by_kernel(ctx, ctx->xfer_down.as<u8>() + ctx->down_off, dev_src, n);
  ctx->reads.push_back({host_dst, ctx->down_off, n});
  ctx->down_off += need;
}

inline void sync_reads(dcdf_ctx* ctx) {
  CK(cudaStreamSynchronize(ctx->stream));
  for (const auto& r : ctx->reads) memcpy(r.dst, ctx->xfer_down.as<u8>() + r.off, r.n);
  ctx->reads.clear();
  ctx->down_off = 0;
  ctx->up_off = 0;
}

}  // namespace dcdf
