// filter_project_tma.cu — the production filter+project kernel: persistent, warp-specialised,
// fed by the TMA engine.
//
//   producer warp : one elected lane issues cp.async.bulk (SASS UBLKCP) copies of the next tiles'
//                   column slices HBM -> shared memory into a ring of stages, completion tracked by
//                   mbarrier transaction counts (no registers, no LSU instructions on the load path)
//   16 consumer   : evaluate the predicate over the staged tile (expression VM, K rows per lane),
//   warps           keep the mask in registers (ballots), later evaluate the projections from the
//                   same staged data and store selected rows at their compacted global position
//   scan warp     : turns the 16 per-warp counts of a tile into global output offsets
//
// Tiles are assigned round-robin: in "wave" `it` CTA c owns tile it*G + c (G = gridDim.x = one
// CTA per SM, all co-resident: cooperative launch).  A chained look-back would serialise on the
// previous tile's owner once per tile, i.e. one L2 round trip of latency per tile per CTA, which
// is longer than the tile's HBM time.  Instead every scan warp publishes its tile's count and then
// GATHERS the counts of all G tiles of its wave in one batch of parallel loads: its own offset is
// base + sum(counts of lower CTAs), and base advances by the wave total — computed redundantly by
// every CTA, so no value is ever forwarded from one wave to the next through memory.
//
// Consumers run the predicate TM_LAG tiles ahead of the projections, so the gather latency of a
// wave is hidden behind useful work; nothing in the CTA executes __syncthreads in the steady state
// (all hand-offs are mbarriers).
//
// Reference path replaced: src/execution/filter.rs:46-110 + src/execution/projection.rs:46-66.
#include "filter_project.cuh"

namespace dfgpu {

constexpr int TM_CWARPS = 16;                     // consumer warps
constexpr int TM_THREADS = (TM_CWARPS + 2) * 32;  // + producer warp + scan warp
constexpr int TM_MAX_STAGES = 8;
constexpr int TM_LAG = 2;    // predicate runs this many tiles ahead of the projection
constexpr int TM_RING = 4;   // slots of the count/offset hand-off rings (> TM_LAG + 1)
constexpr int TM_MAX_GRID = 256;
constexpr int TM_SMEM_BUDGET = 200 * 1024;

// ---- mbarrier / bulk-copy PTX ----------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
#ifdef DF_TRYWAIT_HINT
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, " DF_TRYWAIT_HINT ";\n"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
#endif
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared, completion reported to an mbarrier (TMA engine; UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct TmaShared {
  unsigned long long full[TM_MAX_STAGES];   // producer -> consumers: stage holds the tile
  unsigned long long empty[TM_MAX_STAGES];  // consumers -> producer: stage may be overwritten
  unsigned long long cnt_ready[TM_RING];    // consumers -> scan warp: per-warp counts of a tile are in s_cnt
  unsigned long long pfx_ready[TM_RING];    // scan warp -> consumers: global offsets of a tile are in s_off
  unsigned s_cnt[TM_RING][TM_CWARPS];
  unsigned long long s_off[TM_RING][TM_CWARPS];
};

template <int DEPTH, int K, bool F64ONLY>
__global__ void __launch_bounds__(TM_THREADS, 1) k_filter_project_tma(const __grid_constant__ FPParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int TILE = TM_CWARPS * 32 * K;
  TmaShared& sh = *reinterpret_cast<TmaShared*>(smem_raw);
  unsigned char* stages = smem_raw + 1024;  // stage ring starts 1 KiB in (keeps 128-B alignment)
  const int S = p.nstages;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(&sh.full[s], 1);
      mbar_init(&sh.empty[s], TM_CWARPS);
    }
    for (int i = 0; i < TM_RING; i++) {
      mbar_init(&sh.cnt_ready[i], TM_CWARPS);
      mbar_init(&sh.pfx_ready[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int first = blockIdx.x, step = gridDim.x;

  if (warp == TM_CWARPS) {
    // ================================ producer warp =============================================
    int it = 0;
    for (int tile = first; tile < p.ntiles; tile += step, ++it) {
      const int s = it % S;
      mbar_wait(&sh.empty[s], ((it / S) & 1) ^ 1);  // first pass over the ring returns immediately
      unsigned char* dst = stages + (size_t)s * p.stage_bytes;
      const long long row0 = (long long)tile * TILE;
      const long long left = p.nrows - row0;
      if (left >= TILE) {
        if (lane == 0) {
          unsigned total = 0;
          for (int c = 0; c < p.ps.ncols; c++) total += (unsigned)(TILE * p.col_w[c]);
          mbar_arrive_expect_tx(&sh.full[s], total);
          for (int c = 0; c < p.ps.ncols; c++)
            tma_load_1d(dst + p.col_off[c], (const unsigned char*)p.ps.cols[c].ptr + row0 * p.col_w[c], (unsigned)(TILE * p.col_w[c]),
                        &sh.full[s]);
        }
      } else {
        // ragged last tile: sizes need not be 16-byte multiples, so the warp copies it by hand
        for (int c = 0; c < p.ps.ncols; c++) {
          const unsigned char* src = (const unsigned char*)p.ps.cols[c].ptr + row0 * p.col_w[c];
          const long long nb = left * p.col_w[c];
          for (long long b = lane; b < nb; b += 32) dst[p.col_off[c] + b] = src[b];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.full[s]);
      }
    }
  } else if (warp == TM_CWARPS + 1) {
    // ================================ scan warp =================================================
    unsigned long long base = 0;  // selected rows in all earlier waves (identical in every CTA)
    int it = 0;
    for (int tile = first; tile < p.ntiles; tile += step, ++it) {
      const int b = it % TM_RING;
      mbar_wait(&sh.cnt_ready[b], (it / TM_RING) & 1);
      const unsigned c = lane < TM_CWARPS ? sh.s_cnt[b][lane] : 0u;
      unsigned incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long prefix;
      if (!p.has_pred) {
        prefix = (unsigned long long)tile * TILE;  // nothing is dropped: positions are known
      } else {
        if (lane == 0) st_relaxed(&p.tile_status[tile], ST_AGG | total);
        // gather the counts of every tile of this wave: all loads issued before the first is used
        const long long wave0 = (long long)it * step;
        unsigned long long sv[TM_MAX_GRID / 32];
#pragma unroll
        for (int w = 0; w < TM_MAX_GRID / 32; w++) {
          const int j = w * 32 + lane;
          const long long idx = wave0 + j;
          sv[w] = (w * 32 < step && j < step && idx < p.ntiles) ? ld_relaxed(&p.tile_status[idx]) : ST_AGG;
        }
        unsigned long long before = 0, wave_total = 0;
#pragma unroll
        for (int w = 0; w < TM_MAX_GRID / 32; w++) {
          if (w * 32 < step) {
            const int j = w * 32 + lane;
            const long long idx = wave0 + j;
            while (__any_sync(0xffffffffu, (sv[w] >> 62) == 0)) {
              if ((sv[w] >> 62) == 0) sv[w] = ld_relaxed(&p.tile_status[idx]);
            }
            const unsigned long long v = sv[w] & ST_MASK;
            wave_total += v;
            if (j < (int)blockIdx.x) before += v;
          }
        }
        before = warp_sum64(before);
        wave_total = warp_sum64(wave_total);
        prefix = base + before;
        base += wave_total;
      }
      if (lane < TM_CWARPS) sh.s_off[b][lane] = prefix + (incl - c);
      if (lane == 0 && tile == p.ntiles - 1) *p.out_count = prefix + total;
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.pfx_ready[b]);
    }
  } else {
    // ================================ consumer warps ============================================
    const unsigned lt_mask = (1u << lane) - 1u;
    bool bad = false;

    // predicate of local iteration `it` -> flag bits (bit k = row warp*32*K + k*32 + lane of the tile)
    auto phase1 = [&](int it, int tile) -> unsigned {
      const int s = it % S;
      mbar_wait(&sh.full[s], (it / S) & 1);
      StagedTile<K> src;
      src.stage = stages + (size_t)s * p.stage_bytes;
      src.col_off = p.col_off;
      src.lrow0 = warp * 32 * K + lane;
      const long long row0 = (long long)tile * TILE + src.lrow0;
      src.valid = 0;
#pragma unroll
      for (int k = 0; k < K; k++)
        if (row0 + k * 32 < p.nrows) src.valid |= 1u << k;
      unsigned flags = src.valid;
      if (p.has_pred && p.pred_fast.kind >= 2) {
        // fast shape: one Float64 comparison, operands straight from the staged tile
        const double* A = (const double*)(src.stage + p.col_off[p.pred_fast.a]) + src.lrow0;
        const bool bcol = p.pred_fast.kind == 2;
        const double* B = bcol ? (const double*)(src.stage + p.col_off[p.pred_fast.b]) + src.lrow0 : A;
        const double imm = p.pred_fast.imm;
        flags = 0;
#define DF_CMP(OPR)                                                        \
  if (bcol) {                                                              \
    _Pragma("unroll") for (int k = 0; k < K; k++) flags |= (unsigned)(A[k * 32] OPR B[k * 32]) << k; \
  } else {                                                                 \
    _Pragma("unroll") for (int k = 0; k < K; k++) flags |= (unsigned)(A[k * 32] OPR imm) << k;       \
  }
        switch (p.pred_fast.op) {
          case V_EQ: DF_CMP(==) break;
          case V_NE: DF_CMP(!=) break;
          case V_LT: DF_CMP(<) break;
          case V_LE: DF_CMP(<=) break;
          case V_GT: DF_CMP(>) break;
          default: DF_CMP(>=) break;
        }
#undef DF_CMP
        flags &= src.valid;
      } else if (p.has_pred) {
        unsigned long long v[K];
        const unsigned b = eval_program<DEPTH, K, F64ONLY>(p.ps, 0, src, v);
        bad = bad || (b != 0);
        flags = 0;
#pragma unroll
        for (int k = 0; k < K; k++) flags |= (unsigned)(v[k] & 1ull) << k;
        flags &= src.valid;
      }
      unsigned cnt = 0;
#pragma unroll
      for (int k = 0; k < K; k++) cnt += __popc(__ballot_sync(0xffffffffu, (flags >> k) & 1u));
      if (lane == 0) {
        sh.s_cnt[it % TM_RING][warp] = cnt;
        mbar_arrive(&sh.cnt_ready[it % TM_RING]);
      }
      return flags;
    };

    // projections of local iteration `it`: selected rows go to their compacted global position
    auto phase2 = [&](int it, int tile, unsigned flags) {
      const int s = it % S;
      mbar_wait(&sh.pfx_ready[it % TM_RING], (it / TM_RING) & 1);
      const unsigned long long base = sh.s_off[it % TM_RING][warp];
      StagedTile<K> src;
      src.stage = stages + (size_t)s * p.stage_bytes;
      src.col_off = p.col_off;
      src.lrow0 = warp * 32 * K + lane;
      src.valid = flags;  // a zero divisor only matters on rows that survive the filter
      for (int q = 0; q < p.nproj; q++) {
        const int prog = q + p.has_pred;
        unsigned long long v[K];
        const FastOp& fo = p.proj_fast[q];
        if (fo.kind == 1) {
          const unsigned long long* A = (const unsigned long long*)(src.stage + p.col_off[fo.a]) + src.lrow0;
#pragma unroll
          for (int k = 0; k < K; k++) v[k] = A[k * 32];
        } else if (fo.kind >= 2) {
          const double* A = (const double*)(src.stage + p.col_off[fo.a]) + src.lrow0;
          const bool bcol = fo.kind == 2;
          const double* B = bcol ? (const double*)(src.stage + p.col_off[fo.b]) + src.lrow0 : A;
          const double imm = fo.imm;
          double y[K];
#pragma unroll
          for (int k = 0; k < K; k++) y[k] = bcol ? B[k * 32] : imm;
          switch (fo.op) {
            case V_ADD:
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = d2u(A[k * 32] + y[k]);
              break;
            case V_SUB:
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = d2u(A[k * 32] - y[k]);
              break;
            case V_MUL:
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = d2u(A[k * 32] * y[k]);
              break;
            default:
#pragma unroll
              for (int k = 0; k < K; k++) {
                if (y[k] == 0.0 && ((flags >> k) & 1u)) bad = true;  // DivideByZero on a surviving row
                v[k] = d2u(A[k * 32] / y[k]);
              }
              break;
          }
        } else {
          const unsigned b = eval_program<DEPTH, K, F64ONLY>(p.ps, prog, src, v);
          bad = bad || (b != 0);
        }
        const int odt = p.ps.out_dtype[prog];
        void* o = p.out[q];
        unsigned long long run = base;
#pragma unroll
        for (int k = 0; k < K; k++) {
          const bool f = (flags >> k) & 1u;
          const unsigned m = __ballot_sync(0xffffffffu, f);
          if (f) {
            const long long idx = (long long)(run + __popc(m & lt_mask));
            if (F64ONLY || fo.kind) ((unsigned long long*)o)[idx] = v[k];
            else store_elem(o, odt, idx, v[k]);
          }
          run += __popc(m);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty[s]);  // this warp is done reading the stage
    };

    // software pipeline: predicate of tile it, then projections of tile it - TM_LAG
    static_assert(TM_LAG == 2, "the flag registers below are written out for a lag of 2");
    int it = 0;
    unsigned f1 = 0, f2 = 0;  // flags of tiles it-1, it-2
    int t1 = -1, t2 = -1;
    for (int tile = first; tile < p.ntiles; tile += step, ++it) {
      const unsigned f0 = phase1(it, tile);
      if (it >= TM_LAG) phase2(it - TM_LAG, t2, f2);
      f2 = f1; t2 = t1;
      f1 = f0; t1 = tile;
    }
    // drain
    if (it >= 2) phase2(it - 2, t2, f2);
    if (it >= 1) phase2(it - 1, t1, f1);
    if (bad) *p.err_flag = 1u;
  }
}

template <int DEPTH, int K, bool F64ONLY>
static void launch_one(dfgpu_ctx* ctx, const FPParams& p, size_t smem) {
  auto kern = k_filter_project_tma<DEPTH, K, F64ONLY>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    DF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BUDGET + 2048));
    configured = true;
  }
  long long grid = std::min(ctx->sm_count, TM_MAX_GRID);  // one persistent CTA per SM
  if (grid > p.ntiles) grid = p.ntiles;
  const int ps = ctx->prof_begin();
  // cooperative launch: the wave-synchronous scan needs every CTA of the grid resident at once
  void* args[] = {(void*)&p};
  DF_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(TM_THREADS), args, smem, ctx->stream));
  ctx->prof_end(ps);
  ctx->launches++;
}

template <int DEPTH, int K>
static void launch_k(dfgpu_ctx* ctx, const FPParams& p, size_t smem) {
  if (p.ps.f64_only) launch_one<DEPTH, K, true>(ctx, p, smem);
  else launch_one<DEPTH, K, false>(ctx, p, smem);
}

bool launch_fp_tma(dfgpu_ctx* ctx, FPParams& p) {
  if (p.ps.max_depth > 4 || p.ps.ncols < 1) return false;
  int row_bytes = 0;
  for (int c = 0; c < p.ps.ncols; c++) {
    p.col_w[c] = dtype_width(p.ps.cols[c].dtype);
    if (p.col_w[c] <= 0) return false;
    row_bytes += p.col_w[c];
  }
  // rows per lane K in {8,4,2}: the biggest tile that still leaves TM_LAG + 3 stages in shared
  // memory (TM_LAG + 1 tiles are held by the consumers, the rest is prefetch depth)
  int K = 0;
  for (int k : {8, 4, 2}) {
    if (k == 8 && p.ps.max_depth > 2) continue;  // deep register stacks spill at 8 rows per lane
    const long long tile = (long long)TM_CWARPS * 32 * k;
    if (tile * row_bytes * (TM_LAG + 3) <= TM_SMEM_BUDGET) { K = k; break; }
  }
  if (!K) return false;
  const int tile = TM_CWARPS * 32 * K;
  int off = 0;
  for (int c = 0; c < p.ps.ncols; c++) {
    p.col_off[c] = off;
    off += tile * p.col_w[c];  // tile is a multiple of 512 rows: every column slice stays 128-B aligned
  }
  p.stage_bytes = off;
  p.nstages = std::min(TM_MAX_STAGES, TM_SMEM_BUDGET / p.stage_bytes);
  p.ntiles = int((p.nrows + tile - 1) / tile);
  const size_t smem = 1024 + (size_t)p.nstages * p.stage_bytes;
  const int d = p.ps.max_depth;
  if (K == 8) launch_k<2, 8>(ctx, p, smem);
  else if (K == 4) { if (d <= 2) launch_k<2, 4>(ctx, p, smem); else launch_k<4, 4>(ctx, p, smem); }
  else { if (d <= 2) launch_k<2, 2>(ctx, p, smem); else launch_k<4, 2>(ctx, p, smem); }
  return true;
}

}  // namespace dfgpu
