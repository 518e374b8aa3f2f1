// utf8_gather.cu — order-preserving gather of a Utf8 (arrow 0.12 BinaryArray: i32 offsets + bytes)
// column by a list of selected row numbers.  Replaces the Utf8 arm of `fn filter`
// (src/execution/filter.rs:93-103: per-row String allocation + BinaryArray::from(Vec<&str>)).
// The row numbers come out of the fused filter kernel as one more projected column (V_PUSH_ROWID).
#include "common.cuh"

namespace dfgpu {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

// lengths of the selected strings, written to out[i + 1] (out[0] = 0 is set by the host)
__global__ void k_utf8_lengths(const unsigned long long* __restrict__ idx, const int* __restrict__ off, long long n, int* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = (long long)idx[i];
    out[i + 1] = off[r + 1] - off[r];
  }
}

// in-place inclusive scan, three launches: per-block scan + block totals, scan of totals, add-back
__global__ void __launch_bounds__(SC_THREADS) k_scan_block(int* __restrict__ a, long long n, long long* __restrict__ sums) {
  __shared__ long long s_warp[SC_THREADS / 32];
  const long long base = (long long)blockIdx.x * SC_TILE + (long long)threadIdx.x * SC_ITEMS;
  long long v[SC_ITEMS];
  long long run = 0;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++) {
    v[k] = base + k < n ? (long long)a[base + k] : 0;
    run += v[k];
    v[k] = run;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  long long wbase = 0;
  for (int w = 0; w < warp; w++) wbase += s_warp[w];
  const long long excl = wbase + incl - run;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; k++)
    if (base + k < n) a[base + k] = (int)(v[k] + excl);  // block-local; the add-back finishes it
  if (threadIdx.x == SC_THREADS - 1) sums[blockIdx.x] = excl + run;
}
__global__ void k_scan_sums(long long* sums, long long nblocks) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long run = 0;
    for (long long b = 0; b < nblocks; b++) {
      run += sums[b];
      sums[b] = run;
    }
  }
}
__global__ void __launch_bounds__(SC_THREADS) k_scan_add(int* __restrict__ a, long long n, const long long* __restrict__ sums) {
  if (blockIdx.x == 0) return;
  const long long add = sums[blockIdx.x - 1];
  const long long base = (long long)blockIdx.x * SC_TILE;
  for (int k = threadIdx.x; k < SC_TILE; k += SC_THREADS)
    if (base + k < n) a[base + k] = (int)((long long)a[base + k] + add);
}

// one warp per selected row: copy its bytes
__global__ void k_utf8_copy(const unsigned long long* __restrict__ idx, const int* __restrict__ off, const unsigned char* __restrict__ bytes,
                            long long n, const int* __restrict__ new_off, unsigned char* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp0; i < n; i += nwarps) {
    const long long r = (long long)idx[i];
    const int s = off[r], len = off[r + 1] - s, d = new_off[i];
    for (int b = lane; b < len; b += 32) out[d + b] = bytes[s + b];
  }
}

// gather `src` (Utf8) by `d_idx[0..nsel)` into `out`.  Synchronises the stream once (byte count).
void gather_utf8(dfgpu_ctx* ctx, const DevColumn& src, const unsigned long long* d_idx, long long nsel, DevColumn* out) {
  out->dtype = DFGPU_UTF8;
  out->offsets = (int32_t*)ctx->alloc(size_t(nsel + 1) * 4);
  DF_CUDA(cudaMemsetAsync(out->offsets, 0, 4, ctx->stream));
  long long total = 0;
  if (nsel > 0) {
    const int grid = (int)std::min<long long>((nsel + 255) / 256, (long long)ctx->sm_count * 8);
    k_utf8_lengths<<<grid, 256, 0, ctx->stream>>>(d_idx, src.offsets, nsel, out->offsets);
    DF_CUDA(cudaGetLastError());
    const long long nblocks = (nsel + SC_TILE - 1) / SC_TILE;
    long long* sums = (long long*)ctx->alloc(size_t(nblocks) * 8);
    k_scan_block<<<(unsigned)nblocks, SC_THREADS, 0, ctx->stream>>>(out->offsets + 1, nsel, sums);
    DF_CUDA(cudaGetLastError());
    k_scan_sums<<<1, 32, 0, ctx->stream>>>(sums, nblocks);
    DF_CUDA(cudaGetLastError());
    k_scan_add<<<(unsigned)nblocks, SC_THREADS, 0, ctx->stream>>>(out->offsets + 1, nsel, sums);
    DF_CUDA(cudaGetLastError());
    ctx->launches += 4;
    DF_CUDA(cudaMemcpyAsync(ctx->h_scratch + 24, sums + (nblocks - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
    DF_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->free(sums);
    total = (long long)ctx->h_scratch[24];
    if (total >= (1ll << 31)) fail(DFGPU_ERR_NOT_IMPLEMENTED, "Utf8 output larger than 2 GiB (i32 offsets)");
  }
  out->values_bytes = size_t(total);
  out->values = ctx->alloc(size_t(total > 0 ? total : 1));
  if (total > 0) {
    const int grid = (int)std::min<long long>((nsel * 32 + 255) / 256, (long long)ctx->sm_count * 16);
    k_utf8_copy<<<grid, 256, 0, ctx->stream>>>(d_idx, src.offsets, (const unsigned char*)src.values, nsel, out->offsets, (unsigned char*)out->values);
    DF_CUDA(cudaGetLastError());
    ctx->launches++;
  }
}

}  // namespace dfgpu
