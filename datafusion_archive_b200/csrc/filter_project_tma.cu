// filter_project_tma.cu — the production filter+project kernel: persistent, warp-specialised,
// fed by the TMA engine, with the predicate running far ahead of the projection.
//
// Roles inside one CTA (one CTA per SM, cooperative launch, tiles assigned round-robin: in "wave"
// `it` CTA c owns tile it*G + c):
//
//   producer A (1 warp) : one elected lane streams the PREDICATE columns of tile it into ring A with
//                         cp.async.bulk (SASS UBLKCP), completion tracked by mbarrier tx counts
//   producer B (1 warp) : streams the PROJECTION columns of tile it-LAG into ring B the same way —
//                         a re-read of bytes fetched LAG tiles earlier, served by the 126 MB L2
//   16 consumer warps   : phase 1 = predicate of tile it over ring A -> K flag bits per lane, kept in
//                         a 64-bit shift register; phase 2 = projections of tile it-LAG over ring B,
//                         selected rows stored at their compacted global position
//   2 scan warps        : turn the per-warp counts of a tile into global output offsets, 4 waves per
//                         warp at a time (8 cross-CTA gathers in flight per SM)
//
// Why the lag: order-preserving compaction needs, per tile, the number of selected rows in ALL
// earlier tiles.  Under a bandwidth-saturating stream every dependent global round trip costs
// microseconds, far longer than a tile's HBM time (~0.5 us), and a tile cannot wait in shared
// memory that long (bandwidth x latency exceeds the SM's storage).  So the predicate pass, which
// only produces 1 bit per row, runs LAG tiles ahead; by the time the projection pass reaches a
// tile its offset has long been resolved, and the tile's bytes come back from L2, not HBM.
//
// The lag must be >= TM_BATCH - 1: a scan warp waits for the counts of a whole batch of waves, and the
// consumers only reach the projection pass of wave w after the predicate pass of wave w + LAG.
//
// Offsets: every scan warp publishes its tile's count, then GATHERS the counts of all G tiles of
// its wave with one batch of parallel loads: offset = base + sum(counts of lower CTAs); base
// advances by the wave total, computed redundantly by every CTA (nothing is forwarded between
// waves through memory).  The scan warps take batches of waves round-robin; the
// running base is handed from wave to wave through shared memory.
//
// Nothing in the CTA executes __syncthreads in the steady state; all hand-offs are mbarriers.
// Reference path replaced: src/execution/filter.rs:46-110 + src/execution/projection.rs:46-66.
#include "filter_project.cuh"

namespace dfgpu {

constexpr int TM_CWARPS = 16;  // consumer warps
constexpr int TM_SWARPS = 2;   // scan warps (2 x TM_BATCH gathers in flight; 20 warps = 5 per SM sub-partition leave 96 registers per thread; a third scan warp was measured: 6 warps on one sub-partition cap the kernel at 80 registers and C2 went from 0.260 to 0.278 ms, profiles/r02_history.md)
#ifndef DF_TM_BATCH
#define DF_TM_BATCH 4
#endif
constexpr int TM_BATCH = DF_TM_BATCH;  // waves per scan-warp batch.  Measured with the lean consumer loop (profiles/r02x_sweep_fp_scan_variants.txt): 2 -> C2 0.298 ms, 3 -> 0.260, 4 -> 0.239; issuing a wave's gather right behind its own publish: 0.33-0.41; one wave per step with the gather consumed a step later (profiles/r02z_sweep_fp_scan_pipe2.txt): 0.38 at every lag (statuses read right after the publish are stale and the re-poll is serial)
constexpr int TM_WARPS = TM_CWARPS + 2 + TM_SWARPS;
constexpr int TM_THREADS = TM_WARPS * 32;
constexpr int TM_MAX_STAGES = 8;
constexpr int TM_RING = 32;       // slots of the count/offset hand-off rings (> max lag + 1)
constexpr int TM_MAX_LAG = 24;
constexpr int TM_MAX_GRID = 160;  // CTAs (= SMs) the wave gather is written for (B200: 148)
constexpr int TM_HDR_BYTES = 8192;
constexpr int TM_SMEM_BUDGET = 200 * 1024;
// upper bound of one hardware suspension in mbarrier.try_wait: a waiting warp sleeps until the phase
// completes (or this long) instead of re-issuing the poll; ncu showed 22 % of all issued instructions in
// the poll loop with the default (short) limit
constexpr unsigned TM_WAIT_HINT_NS = 4000;

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
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(TM_WAIT_HINT_NS)
      : "memory");
}
// the same on 32-bit shared-window addresses kept in registers (the lean consumer loop: no generic -> shared
// conversion, S2UR SR_CgaCtaId + ULEA, in front of every barrier operation)
__device__ __forceinline__ void mbar_arrive_a(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(TM_WAIT_HINT_NS)
      : "memory");
}
// 1-D bulk async copy global -> shared, completion reported to an mbarrier (TMA engine; UBLKCP),
// with an L2 eviction-priority hint: the predicate stream marks bytes that the projection stream
// will re-read LAG tiles later as evict_last; the projection stream (and bytes read once) use
// evict_first so they do not push the pending re-reads out of L2.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar,
                                            unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

struct TmaShared {
  unsigned long long fullA[TM_MAX_STAGES], emptyA[TM_MAX_STAGES];  // ring A (predicate columns)
  unsigned long long fullB[TM_MAX_STAGES], emptyB[TM_MAX_STAGES];  // ring B (projection columns)
  unsigned long long cnt_ready[TM_RING];   // consumers -> scan warp: per-warp counts of a tile are in s_cnt
  unsigned long long pfx_ready[TM_RING];   // scan warp -> consumers: global offsets of a tile are in s_off
  unsigned long long base_ready[TM_RING];  // scan warp of wave it-1 -> scan warp of wave it
  unsigned long long s_base[TM_RING];
  unsigned long long s_off[TM_RING][TM_CWARPS];
  unsigned s_cnt[TM_RING][TM_CWARPS];
};
static_assert(sizeof(TmaShared) <= TM_HDR_BYTES, "shared header too large");

// One producer warp: stream the `ncols` columns listed in `slots` of every tile this CTA owns into
// a ring of S stages.
__device__ __forceinline__ void producer_loop(const FPParams& p, int tile_rows, const int* col_off, const int* reread_off, unsigned char* ring,
                                              int S, int stage_bytes, unsigned long long* full, unsigned long long* empty, int lane) {
  const unsigned long long keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
  int s = 0;
  unsigned ph = 1;  // first pass over the ring returns immediately
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    mbar_wait(&empty[s], ph);
    unsigned char* dst = ring + (size_t)s * stage_bytes;
    const long long row0 = (long long)tile * tile_rows;
    const long long left = p.nrows - row0;
    if (left >= tile_rows) {
      if (lane == 0) {
        unsigned total = 0;
        for (int c = 0; c < p.ps.ncols; c++)
          if (col_off[c] >= 0) total += (unsigned)(tile_rows * p.col_w[c]);
        mbar_arrive_expect_tx(&full[s], total);
        for (int c = 0; c < p.ps.ncols; c++)
          if (col_off[c] >= 0)
            tma_load_1d(dst + col_off[c], (const unsigned char*)p.ps.cols[c].ptr + row0 * p.col_w[c], (unsigned)(tile_rows * p.col_w[c]),
                        &full[s], (reread_off && reread_off[c] >= 0) ? keep : stream);
      }
    } else {
      // ragged last tile: sizes need not be 16-byte multiples, so the warp copies it by hand
      for (int c = 0; c < p.ps.ncols; c++) {
        if (col_off[c] < 0) continue;
        const unsigned char* src = (const unsigned char*)p.ps.cols[c].ptr + row0 * p.col_w[c];
        const long long nb = left * p.col_w[c];
        for (long long b = lane; b < nb; b += 32) dst[col_off[c] + b] = src[b];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
    if (++s == S) { s = 0; ph ^= 1u; }
  }
}

// one comparison over the K rows of this lane -> K flag bits (operands straight from the staged tile)
template <int K, class T>
__device__ __forceinline__ unsigned cmp_term_t(const FastOp& t, const unsigned char* stage, const int* col_off, int lrow0, T imm) {
  const T* A = (const T*)(stage + col_off[t.a]) + lrow0;
  const bool bcol = t.kind == 2;
  const T* B = bcol ? (const T*)(stage + col_off[t.b]) + lrow0 : A;
  unsigned flags = 0;
#define DF_CMP(OPR)                                                                                    \
  if (bcol) {                                                                                          \
    _Pragma("unroll") for (int k = 0; k < K; k++) flags |= (unsigned)(A[k * 32] OPR B[k * 32]) << k;   \
  } else {                                                                                             \
    _Pragma("unroll") for (int k = 0; k < K; k++) flags |= (unsigned)(A[k * 32] OPR imm) << k;         \
  }
  switch (t.op) {
    case V_EQ: DF_CMP(==) break;
    case V_NE: DF_CMP(!=) break;
    case V_LT: DF_CMP(<) break;
    case V_LE: DF_CMP(<=) break;
    case V_GT: DF_CMP(>) break;
    default: DF_CMP(>=) break;
  }
#undef DF_CMP
  return flags;
}
template <int K, bool F64>
__device__ __forceinline__ unsigned cmp_term(const FastOp& t, const unsigned char* stage, const int* col_off, int lrow0) {
  if (F64) return cmp_term_t<K, double>(t, stage, col_off, lrow0, u2d(t.imm));
  switch (t.ty) {
    case DFGPU_FLOAT64: return cmp_term_t<K, double>(t, stage, col_off, lrow0, u2d(t.imm));
    case DFGPU_INT64: return cmp_term_t<K, long long>(t, stage, col_off, lrow0, (long long)t.imm);
    case DFGPU_UINT64: return cmp_term_t<K, unsigned long long>(t, stage, col_off, lrow0, t.imm);
    case DFGPU_FLOAT32: return cmp_term_t<K, float>(t, stage, col_off, lrow0, u2f(t.imm));
    case DFGPU_INT32: return cmp_term_t<K, int>(t, stage, col_off, lrow0, (int)(long long)t.imm);
    default: return cmp_term_t<K, unsigned>(t, stage, col_off, lrow0, (unsigned)t.imm);
  }
}

// one arithmetic operation over the K rows of this lane (Float64 / Float32 / 64-bit integers)
template <int K, class T>
__device__ __forceinline__ void arith_term_t(const FastOp& t, const unsigned char* stage, const int* col_off, int lrow0, T imm, unsigned flags,
                                             bool& bad, T (&out)[K]) {
  const T* A = (const T*)(stage + col_off[t.a]) + lrow0;
  const bool bcol = t.kind == 2;
  const T* B = bcol ? (const T*)(stage + col_off[t.b]) + lrow0 : A;
  T y[K];
#pragma unroll
  for (int k = 0; k < K; k++) y[k] = bcol ? B[k * 32] : imm;
  switch (t.op) {
    case V_ADD:
#pragma unroll
      for (int k = 0; k < K; k++) out[k] = A[k * 32] + y[k];
      break;
    case V_SUB:
#pragma unroll
      for (int k = 0; k < K; k++) out[k] = A[k * 32] - y[k];
      break;
    case V_MUL:
#pragma unroll
      for (int k = 0; k < K; k++) out[k] = A[k * 32] * y[k];
      break;
    default:  // V_DIV: floating point only (the host does not select integer division as a fast shape)
#pragma unroll
      for (int k = 0; k < K; k++) {
        if (y[k] == T(0) && ((flags >> k) & 1u)) bad = true;  // DivideByZero on a surviving row
        out[k] = A[k * 32] / y[k];
      }
      break;
  }
}

// ---- lean consumer loop ---------------------------------------------------------------------------
// The shapes the headline configurations have (C2: SELECT a WHERE a > c; C3: SELECT a+b, a*b WHERE b < a): ONE
// Float64 comparison as the predicate and one or two projections that copy an 8-byte column or combine Float64
// operands.  ncu on the generic FAST loop (profiles/r02_c2.lines.txt, SASS view in profiles/r02_history.md): 383
// instructions per warp-tile of which ~120 touch rows; the rest re-derives per-tile invariants — indexed
// constant-bank loads of the column offsets behind the term's column index, a jump table on the comparison
// operator, generic -> shared address conversions in front of every mbarrier operation — and its dependent
// latencies (short scoreboard 19 %, no-instruction 7 %, branch resolving 5 % of the consumer samples) are what the
// 4 consumer warps per scheduler cannot hide.  Here the comparison operator and the operand kind are template
// parameters, every offset, pointer and barrier address is computed once before the loop, and the loop body is
// waits + loads + compares + the ballot-compacted store.  Protocol (barriers, rings, scan warps) unchanged.
template <int CMP, class T>
__device__ __forceinline__ bool lean_cmp_t(T a, T b) {
  if (CMP == V_EQ) return a == b;
  if (CMP == V_NE) return a != b;
  if (CMP == V_LT) return a < b;
  if (CMP == V_LE) return a <= b;
  if (CMP == V_GT) return a > b;
  return a >= b;
}
// TY: 0 = Float64, 1 = Int64, 2 = UInt64 (operands as raw 8-byte words)
template <int CMP, int TY>
__device__ __forceinline__ bool lean_cmp(unsigned long long a, unsigned long long b) {
  if (TY == 0) return lean_cmp_t<CMP, double>(u2d(a), u2d(b));
  if (TY == 1) return lean_cmp_t<CMP, long long>((long long)a, (long long)b);
  return lean_cmp_t<CMP, unsigned long long>(a, b);
}

// predicated 8-byte store (the compiler turns `if (selected) out[pos] = v` into a divergent branch per row when the
// value's load can be sunk into it; the compacted store wants @P STG)
__device__ __forceinline__ void st_if(unsigned cond, unsigned long long* dst, unsigned long long v) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.u32 p, %0, 0;\n"
      "@p st.global.b64 [%1], %2;\n"
      "}\n" ::"r"(cond), "l"(dst), "l"(v)
      : "memory");
}

// 8-byte load from a 32-bit shared-window address (ordered with the mbarrier operations around it: volatile + memory)
__device__ __forceinline__ unsigned long long lds64(unsigned addr) {
  unsigned long long v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}

// One row slot of the ballot compaction, fused: p = (flags & bit) != 0; m = ballot(p); pos = run + popc(m & lanes
// below); @p store v at o[pos]; run += popc(m).  One predicate feeds the vote and the store (the C++ form costs a
// shift + and + compare for the vote and an and + compare again for the store).
__device__ __forceinline__ unsigned compact_store(unsigned flags, unsigned bit, unsigned lt_mask, unsigned& run, unsigned long long* o, unsigned long long v) {
  unsigned pos;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 m, t;\n"
      ".reg .b64 a;\n"
      "and.b32 t, %2, %3;\n"
      "setp.ne.u32 p, t, 0;\n"
      "vote.sync.ballot.b32 m, p, 0xffffffff;\n"
      "and.b32 t, m, %4;\n"
      "popc.b32 t, t;\n"
      "add.u32 %1, t, %0;\n"
      "mad.wide.u32 a, %1, 8, %5;\n"
      "@p st.global.b64 [a], %6;\n"
      "popc.b32 m, m;\n"
      "add.u32 %0, %0, m;\n"
      "}\n"
      : "+r"(run), "=&r"(pos)
      : "r"(flags), "r"(bit), "r"(lt_mask), "l"(o), "l"(v)
      : "memory");
  return pos;
}

constexpr int LEAN_MAX_PROJ = 2;

template <int K, int CMP, bool PB, int NP, int TY>
__device__ __forceinline__ void consumer_lean(const FPParams& p, int warp, int lane) {
  extern __shared__ __align__(128) unsigned char smem_raw[];  // the kernel's dynamic shared memory: 32-bit shared-window arithmetic below
  TmaShared& sh = *reinterpret_cast<TmaShared*>(smem_raw);
  constexpr int TILE = TM_CWARPS * 32 * K;
  constexpr unsigned KMASK = (1u << K) - 1u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int first = blockIdx.x, step = gridDim.x;
  const int nloc = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;
  const int SA = p.nstagesA, SB = p.nstagesB, LAG = p.lag;
  const bool single = p.single_ring != 0;
  const int last_it = (p.ntiles - 1 - first) % step == 0 ? (p.ntiles - 1 - first) / step : -1;  // the only ragged tile, if this CTA owns it
  unsigned sh0 = smem_u32(smem_raw);
  asm volatile("" : "+r"(sh0));  // one register for every barrier address: do not re-derive the shared-window base at every use
  constexpr unsigned b_fullA = (unsigned)offsetof(TmaShared, fullA), b_emptyA = (unsigned)offsetof(TmaShared, emptyA);
  constexpr unsigned b_fullB = (unsigned)offsetof(TmaShared, fullB), b_emptyB = (unsigned)offsetof(TmaShared, emptyB);
  constexpr unsigned b_cnt = (unsigned)offsetof(TmaShared, cnt_ready), b_pfx = (unsigned)offsetof(TmaShared, pfx_ready);
  // operand offsets (bytes from the start of shared memory) in stage 0
  const int lrow0 = warp * 32 * K + lane;
  const FastOp& pt = p.pred_fast.term[0];
  const unsigned long long pimm = pt.imm;
  const int ringA_off = TM_HDR_BYTES, stageA = p.stage_bytesA;
  const int ring2_off = single ? TM_HDR_BYTES : TM_HDR_BYTES + SA * stageA;
  const int stage2 = single ? stageA : p.stage_bytesB;
  const unsigned a1 = sh0 + (unsigned)(ringA_off + p.col_offA[pt.a] + lrow0 * 8);  // shared-window addresses: LDS [R + imm], nothing to derive per tile
  const unsigned b1 = PB ? sh0 + (unsigned)(ringA_off + p.col_offA[pt.b] + lrow0 * 8) : a1;
  unsigned a2[NP], b2[NP];
  int kop2[NP];
  unsigned long long imm2[NP];
  unsigned long long* out2[NP];
#pragma unroll
  for (int q = 0; q < NP; q++) {
    const FastOp& fo = p.proj_fast[q];
    // -1 copy; op (+0x100: the right operand is the immediate; +0x200: 64-bit integer arithmetic, two's complement wrap-around)
    kop2[q] = fo.kind == 1 ? -1 : ((fo.kind == 2 ? fo.op : fo.op | 0x100) | (fo.ty == DFGPU_FLOAT64 ? 0 : 0x200));
    imm2[q] = fo.imm;
    a2[q] = sh0 + (unsigned)(ring2_off + p.col_offB[fo.a] + lrow0 * 8);
    b2[q] = fo.kind == 2 ? sh0 + (unsigned)(ring2_off + p.col_offB[fo.b] + lrow0 * 8) : a2[q];
    out2[q] = (unsigned long long*)p.out[q];
  }
  bool bad = false;
  unsigned __int128 fl = 0;  // flag bits of the last LAG+1 tiles, K per tile
  int sa = 0, sb = 0;
  unsigned pha = 0, phb = 0;
  for (int it = 0; it < nloc + LAG; it++) {
    unsigned f0 = 0;
    if (it < nloc) {
      // ---- predicate of tile `it` -> K flag bits, the warp's count to the scan warp
      mbar_wait_a(sh0 + b_fullA + 8u * sa, pha);
      const unsigned A = a1 + (unsigned)(sa * stageA), B = b1 + (unsigned)(sa * stageA);
      unsigned long long x[K], y[K];
#pragma unroll
      for (int k = 0; k < K; k++) x[k] = lds64(A + k * 256);
#pragma unroll
      for (int k = 0; k < K; k++) y[k] = PB ? lds64(B + k * 256) : pimm;
#pragma unroll
      for (int k = 0; k < K; k++) f0 |= (unsigned)lean_cmp<CMP, TY>(x[k], y[k]) << k;
      if (it == last_it) {
        const long long row0 = ((long long)first + (long long)it * step) * TILE + lrow0;
        unsigned valid = 0;
#pragma unroll
        for (int k = 0; k < K; k++)
          if (row0 + k * 32 < p.nrows) valid |= 1u << k;
        f0 &= valid;
      }
      const unsigned cnt = __reduce_add_sync(0xffffffffu, (unsigned)__popc(f0));
      if (lane == 0) {
        if (!single) mbar_arrive_a(sh0 + b_emptyA + 8u * sa);  // this warp is done reading the stage
        sh.s_cnt[it % TM_RING][warp] = cnt;
        mbar_arrive_a(sh0 + b_cnt + 8u * (it % TM_RING));
      }
      if (++sa == SA) { sa = 0; pha ^= 1u; }
    }
    fl = (fl << K) | (unsigned __int128)f0;
    if (it >= LAG) {
      // ---- projections of tile it - LAG: selected rows go to their compacted global position
      const int j = it - LAG;
      const unsigned flags = (unsigned)(fl >> (K * LAG)) & KMASK;
      mbar_wait_a(sh0 + b_pfx + 8u * (j % TM_RING), (j / TM_RING) & 1);
      const unsigned long long base = sh.s_off[j % TM_RING][warp];
      if (!single) mbar_wait_a(sh0 + b_fullB + 8u * sb, phb);
      // projection values of the K rows of this lane, then the ballot compaction: the rank of a selected row inside the
      // warp's slice (row order: k major, lane minor) is computed while the first projection is stored
      unsigned pos[K];
#pragma unroll
      for (int q = 0; q < NP; q++) {
        const unsigned A = a2[q] + (unsigned)(sb * stage2);
        unsigned long long* o = out2[q] + base;
        unsigned long long v[K];
        if (kop2[q] < 0) {
#pragma unroll
          for (int k = 0; k < K; k++) v[k] = lds64(A + k * 256);
        } else {
          const unsigned B = b2[q] + (unsigned)(sb * stage2);
          const bool rimm = (kop2[q] & 0x100) != 0;
          const int op = kop2[q] & 0xff;
          unsigned long long y[K];
#pragma unroll
          for (int k = 0; k < K; k++) v[k] = lds64(A + k * 256);
#pragma unroll
          for (int k = 0; k < K; k++) y[k] = rimm ? imm2[q] : lds64(B + k * 256);
          if (kop2[q] & 0x200) {  // Int64 / UInt64: + - * (the host keeps integer division out of the fast shapes)
            if (op == V_ADD) {
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = v[k] + y[k];
            } else if (op == V_MUL) {
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = v[k] * y[k];
            } else {
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = v[k] - y[k];
            }
          } else if (op == V_ADD) {
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = d2u(u2d(v[k]) + u2d(y[k]));
          } else if (op == V_MUL) {
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = d2u(u2d(v[k]) * u2d(y[k]));
          } else if (op == V_SUB) {
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = d2u(u2d(v[k]) - u2d(y[k]));
          } else {  // V_DIV
#pragma unroll
            for (int k = 0; k < K; k++) {
              if (u2d(y[k]) == 0.0 && (flags & (1u << k))) bad = true;  // DivideByZero on a surviving row
              v[k] = d2u(u2d(v[k]) / u2d(y[k]));
            }
          }
        }
        if (q == 0) {
          unsigned run = 0;
#pragma unroll
          for (int k = 0; k < K; k++) pos[k] = compact_store(flags, 1u << k, lt_mask, run, o, v[k]);
        } else {
#pragma unroll
          for (int k = 0; k < K; k++) st_if(flags & (1u << k), o + pos[k], v[k]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_a(sh0 + (single ? b_emptyA : b_emptyB) + 8u * sb);
      if (++sb == SB) { sb = 0; phb ^= 1u; }
    }
  }
  if (bad) *p.err_flag = 1u;
}

template <int K, int NP, int TY>
__device__ __forceinline__ void consumer_lean_ops(const FPParams& p, int warp, int lane) {
  const int op = p.pred_fast.term[0].op;
  const bool pb = p.pred_fast.term[0].kind == 2;
#define DF_LEAN(OP)                                                \
  case OP:                                                         \
    if (pb) consumer_lean<K, OP, true, NP, TY>(p, warp, lane);     \
    else consumer_lean<K, OP, false, NP, TY>(p, warp, lane);       \
    break;
  switch (op) {
    DF_LEAN(V_EQ)
    DF_LEAN(V_NE)
    DF_LEAN(V_LT)
    DF_LEAN(V_LE)
    DF_LEAN(V_GT)
    default:
      if (pb) consumer_lean<K, V_GE, true, NP, TY>(p, warp, lane);
      else consumer_lean<K, V_GE, false, NP, TY>(p, warp, lane);
      break;
  }
#undef DF_LEAN
}
template <int K, int NP>
__device__ __forceinline__ void consumer_lean_dispatch(const FPParams& p, int warp, int lane) {
  const int ty = p.pred_fast.term[0].ty;
  if (ty == DFGPU_FLOAT64) consumer_lean_ops<K, NP, 0>(p, warp, lane);
  else if (ty == DFGPU_INT64) consumer_lean_ops<K, NP, 1>(p, warp, lane);
  else consumer_lean_ops<K, NP, 2>(p, warp, lane);
}

// FAST: every program of the query is a fast shape, so the interpreter is not even compiled into
// this instantiation (fewer registers, smaller code).  FAST + F64ONLY: additionally every operand is
// Float64, and the per-type dispatch of the fast shapes disappears too (the C2 / C3 kernels).
template <int DEPTH, int K, bool F64ONLY, bool FAST, bool STASH = false, int LEAN = 0>  // LEAN = number of projections of a lean shape
__global__ void __launch_bounds__(TM_THREADS, 1) k_filter_project_tma(const __grid_constant__ FPParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int TILE = TM_CWARPS * 32 * K;
  TmaShared& sh = *reinterpret_cast<TmaShared*>(smem_raw);
  unsigned char* ringA = smem_raw + TM_HDR_BYTES;
  unsigned char* ringB = ringA + (size_t)p.nstagesA * p.stage_bytesA;
  const int SA = p.nstagesA, SB = p.nstagesB;
  const int LAG = p.lag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < TM_MAX_STAGES; s++) {
      mbar_init(&sh.fullA[s], 1);
      mbar_init(&sh.emptyA[s], TM_CWARPS);
      mbar_init(&sh.fullB[s], 1);
      mbar_init(&sh.emptyB[s], TM_CWARPS);
    }
    for (int i = 0; i < TM_RING; i++) {
      mbar_init(&sh.cnt_ready[i], TM_CWARPS);
      mbar_init(&sh.pfx_ready[i], 1);
      mbar_init(&sh.base_ready[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sh.s_base[0] = 0;  // wave 0 starts from offset 0: its hand-off is pre-completed here
    mbar_arrive(&sh.base_ready[0]);
  }
  __syncthreads();

  const int first = blockIdx.x, step = gridDim.x;

  if (warp == TM_CWARPS) {
    // ================================ producer A: predicate columns ============================
    if (p.has_pred) producer_loop(p, TILE, p.col_offA, STASH ? nullptr : p.col_offB, ringA, SA, p.stage_bytesA, sh.fullA, sh.emptyA, lane);
  } else if (warp == TM_CWARPS + 1) {
    // ================================ producer B: projection columns ===========================
    // Runs as far ahead as ring B allows; the consumers reach these tiles LAG iterations after the
    // predicate pass touched the same rows, so the bytes are L2 hits.
    if (!STASH && !p.single_ring) producer_loop(p, TILE, p.col_offB, nullptr, ringB, SB, p.stage_bytesB, sh.fullB, sh.emptyB, lane);
  } else if (warp >= TM_CWARPS + 2) {
    // ================================ scan warps ================================================
    if (!p.has_pred) return;  // nothing is dropped: output positions are the row numbers
    if (STASH && p.noscan) return;
    const int sw = warp - (TM_CWARPS + 2);
    int nloc = 0;
    for (int tile = first; tile < p.ntiles; tile += step) nloc++;
    // Each scan warp owns batches of TM_BATCH consecutive waves (batch j -> warp j % TM_SWARPS), so
    // TM_SWARPS * TM_BATCH gathers are in flight per CTA: one gather is a full L2 round trip under
    // load, several times longer than a tile.
    for (int w0 = sw * TM_BATCH; w0 < nloc; w0 += TM_SWARPS * TM_BATCH) {
      const int nb = min(TM_BATCH, nloc - w0);
      unsigned excl[TM_BATCH];
      unsigned long long total[TM_BATCH];
      // 1. per-tile counts -> exclusive per-warp offsets, publish the tile totals
      unsigned long long sv[TM_BATCH][TM_MAX_GRID / 32];
#pragma unroll
      for (int i = 0; i < TM_BATCH; i++) {
        excl[i] = 0;
        total[i] = 0;
        if (i < nb) {
          const int it = w0 + i, b = it % TM_RING;
          mbar_wait(&sh.cnt_ready[b], (it / TM_RING) & 1);
          const unsigned c = lane < TM_CWARPS ? sh.s_cnt[b][lane] : 0u;
          unsigned incl = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          excl[i] = incl - c;
          total[i] = __shfl_sync(0xffffffffu, incl, 31);
          if (lane == 0) st_relaxed(&p.tile_status[first + it * step], ST_AGG | total[i]);
        }
      }
      // 2. gather the counts of every tile of these waves: all loads issued before any is used
#pragma unroll
      for (int i = 0; i < TM_BATCH; i++) {
        const long long wave0 = (long long)(w0 + i) * step;
#pragma unroll
        for (int w = 0; w < TM_MAX_GRID / 32; w++) {
          const int j = w * 32 + lane;
          const long long idx = wave0 + j;
          sv[i][w] = (i < nb && w * 32 < step && j < step && idx < p.ntiles) ? ld_relaxed(&p.tile_status[idx]) : ST_AGG;
        }
      }
      unsigned long long before[TM_BATCH], wave_total[TM_BATCH];
#pragma unroll
      for (int i = 0; i < TM_BATCH; i++) {
        const long long wave0 = (long long)(w0 + i) * step;
        unsigned long long bf = 0, wt = 0;
#pragma unroll
        for (int w = 0; w < TM_MAX_GRID / 32; w++) {
          if (i < nb && w * 32 < step) {
            const int j = w * 32 + lane;
            const long long idx = wave0 + j;
            while (__any_sync(0xffffffffu, (sv[i][w] >> 62) == 0)) {
              if ((sv[i][w] >> 62) == 0) sv[i][w] = ld_relaxed(&p.tile_status[idx]);
            }
            const unsigned long long v = sv[i][w] & ST_MASK;
            wt += v;
            if (j < (int)blockIdx.x) bf += v;
          }
        }
        before[i] = warp_sum64(bf);
        wave_total[i] = warp_sum64(wt);
      }
      // 3. running base: handed from the scan warp of the previous batch through shared memory
      const int hb = (w0 / TM_BATCH) % TM_RING;
      mbar_wait(&sh.base_ready[hb], ((w0 / TM_BATCH) / TM_RING) & 1);
      unsigned long long base = sh.s_base[hb];
#pragma unroll
      for (int i = 0; i < TM_BATCH; i++) {
        if (i < nb) {
          const int it = w0 + i, b = it % TM_RING;
          if (lane < TM_CWARPS) sh.s_off[b][lane] = base + before[i] + excl[i];
          if (lane == 0 && first + it * step == p.ntiles - 1) *p.out_count = base + before[i] + total[i];
          base += wave_total[i];
        }
      }
      __syncwarp();
      if (lane == 0) {
        const int nh = (w0 / TM_BATCH + 1) % TM_RING;
        sh.s_base[nh] = base;
        mbar_arrive(&sh.base_ready[nh]);
        for (int i = 0; i < nb; i++) mbar_arrive(&sh.pfx_ready[(w0 + i) % TM_RING]);
      }
    }
  } else if constexpr (LEAN > 0) {
    // ================================ consumer warps, lean shapes ================================
    consumer_lean_dispatch<K, LEAN>(p, warp, lane);
  } else {
    // ================================ consumer warps ============================================
    const unsigned lt_mask = (1u << lane) - 1u;
    constexpr unsigned KMASK = (1u << K) - 1u;
    bool bad = false;

    // predicate of local iteration `it` -> K flag bits (bit k = row warp*32*K + k*32 + lane of the tile)
    auto phase1 = [&](int it, int tile, int s, unsigned ph) -> unsigned {
      mbar_wait(&sh.fullA[s], ph);
      StagedTile<K> src;
      src.stage = ringA + (size_t)s * p.stage_bytesA;
      src.col_off = p.col_offA;
      src.lrow0 = warp * 32 * K + lane;
      src.row0 = (long long)tile * TILE + src.lrow0;
      src.valid = KMASK;
      if (tile == p.ntiles - 1) {  // only the last tile can be ragged
        const long long row0 = src.row0;
        src.valid = 0;
#pragma unroll
        for (int k = 0; k < K; k++)
          if (row0 + k * 32 < p.nrows) src.valid |= 1u << k;
      }
      unsigned flags;
      if (FAST || p.pred_fast.nterms > 0) {
        // fast shape: Float64 comparisons straight from the staged tile, joined by AND / OR
        flags = cmp_term<K, FAST && F64ONLY>(p.pred_fast.term[0], src.stage, p.col_offA, src.lrow0);
        for (int t = 1; t < p.pred_fast.nterms; t++) {
          const unsigned ft = cmp_term<K, FAST && F64ONLY>(p.pred_fast.term[t], src.stage, p.col_offA, src.lrow0);
          flags = p.pred_fast.conn[t] ? (flags | ft) : (flags & ft);
        }
      } else if constexpr (!FAST) {
        unsigned long long v[K];
        const unsigned b = eval_program<DEPTH, K, F64ONLY>(p.ps, 0, src, v);
        bad = bad || (b != 0);
        flags = 0;
#pragma unroll
        for (int k = 0; k < K; k++) flags |= (unsigned)(v[k] & 1ull) << k;
      }
      flags &= src.valid;
      // rows this warp selected in the tile: one population count per lane, one warp reduction (REDUX)
      unsigned cnt = 0;
      if (p.count_ballot) {  // A/B switch (DFGPU_FP_COUNT=ballot): the pre-REDUX form
#pragma unroll
        for (int k = 0; k < K; k++) cnt += __popc(__ballot_sync(0xffffffffu, (flags >> k) & 1u));
        __syncwarp();
      } else {
        cnt = __reduce_add_sync(0xffffffffu, (unsigned)__popc(flags));
      }
      if (lane == 0) {
        if (!p.single_ring) mbar_arrive(&sh.emptyA[s]);  // this warp is done reading the stage
        sh.s_cnt[it % TM_RING][warp] = cnt;
        mbar_arrive(&sh.cnt_ready[it % TM_RING]);
      }
      return flags;
    };

    // projections of local iteration `it`: selected rows go to their compacted global position
    auto phase2 = [&](int it, int tile, unsigned flags, int s, unsigned ph) {
      unsigned long long base;
      if (p.has_pred) {
        mbar_wait(&sh.pfx_ready[it % TM_RING], (it / TM_RING) & 1);
        base = sh.s_off[it % TM_RING][warp];
      } else {
        base = (unsigned long long)tile * TILE + (unsigned long long)warp * 32 * K;
      }
      StagedTile<K> src;
      if (p.single_ring) {
        // the tile is still resident in ring A (held since the predicate pass): no second load
        src.stage = ringA + (size_t)s * p.stage_bytesA;
      } else {
        mbar_wait(&sh.fullB[s], ph);
        src.stage = ringB + (size_t)s * p.stage_bytesB;
      }
      src.col_off = p.col_offB;
      src.lrow0 = warp * 32 * K + lane;
      src.row0 = (long long)tile * TILE + src.lrow0;
      src.valid = flags;  // a zero divisor only matters on rows that survive the filter
      for (int q = 0; q < p.nproj; q++) {
        const int prog = q + p.has_pred;
        unsigned long long v[K];
        const FastOp& fo = p.proj_fast[q];
        if (fo.kind == 1 || (FAST && fo.kind < 2)) {
          if (FAST && F64ONLY) {
            const unsigned long long* A = (const unsigned long long*)(src.stage + p.col_offB[fo.a]) + src.lrow0;
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = A[k * 32];
          } else {
            src.load_rows(p.ps, fo.a, v);
          }
        } else if (fo.kind >= 2) {
          switch ((FAST && F64ONLY) ? (int)DFGPU_FLOAT64 : fo.ty) {
            case DFGPU_FLOAT64: {
              double o[K];
              arith_term_t<K, double>(fo, src.stage, p.col_offB, src.lrow0, u2d(fo.imm), flags, bad, o);
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = d2u(o[k]);
              break;
            }
            case DFGPU_FLOAT32: {
              float o[K];
              arith_term_t<K, float>(fo, src.stage, p.col_offB, src.lrow0, u2f(fo.imm), flags, bad, o);
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = f2u(o[k]);
              break;
            }
            default: {  // Int64 / UInt64: two's complement wrap-around is the natural 64-bit result
              unsigned long long o[K];
              arith_term_t<K, unsigned long long>(fo, src.stage, p.col_offB, src.lrow0, fo.imm, flags, bad, o);
#pragma unroll
              for (int k = 0; k < K; k++) v[k] = o[k];
              break;
            }
          }
        } else if constexpr (!FAST) {
          const unsigned b = eval_program<DEPTH, K, F64ONLY>(p.ps, prog, src, v);
          bad = bad || (b != 0);
        }
        const int odt = p.ps.out_dtype[prog];
        const bool wide = (FAST && F64ONLY) || dtype_width_dev(odt) == 8;
        // warp-local base pointer once (64-bit), then 32-bit running offsets
        unsigned char* o = (unsigned char*)p.out[q] + base * (unsigned long long)(wide ? 8 : dtype_width_dev(odt));
        // compacted store; the element-width dispatch is warp-uniform and hoisted out of the row loop
#define DF_STORE_LOOP(TYPE)                                                        \
  {                                                                                \
    unsigned run = 0;                                                              \
    _Pragma("unroll") for (int k = 0; k < K; k++) {                                \
      const bool f = (flags >> k) & 1u;                                            \
      const unsigned m = __ballot_sync(0xffffffffu, f);                            \
      if (f) ((TYPE*)o)[run + __popc(m & lt_mask)] = (TYPE)v[k];                   \
      run += __popc(m);                                                            \
    }                                                                              \
  }
        if (wide) DF_STORE_LOOP(unsigned long long)
        else switch (dtype_width_dev(odt)) {
          case 4: DF_STORE_LOOP(unsigned) break;
          case 2: DF_STORE_LOOP(unsigned short) break;
          case 1: DF_STORE_LOOP(unsigned char) break;
          default: DF_STORE_LOOP(unsigned long long) break;
        }
#undef DF_STORE_LOOP
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(p.single_ring ? &sh.emptyA[s] : &sh.emptyB[s]);
    };

    // ---- stash mode (FAST shapes over 8-byte columns) -----------------------------------------------------
    // pass 1 of local iteration `it`: predicate, projections, compaction of the selected values into this
    // tile's slab slot at WARP-LOCAL positions (no global offset needed), stage released at once
    auto stash1 = [&](int it, int tile, int s, unsigned ph) {
      mbar_wait(&sh.fullA[s], ph);
      const unsigned char* stage = ringA + (size_t)s * p.stage_bytesA;
      const int lrow0 = warp * 32 * K + lane;
      unsigned valid = KMASK;
      if (tile == p.ntiles - 1) {  // only the last tile can be ragged
        const long long row0 = (long long)tile * TILE + lrow0;
        valid = 0;
#pragma unroll
        for (int k = 0; k < K; k++)
          if (row0 + k * 32 < p.nrows) valid |= 1u << k;
      }
      unsigned flags = cmp_term<K, FAST && F64ONLY>(p.pred_fast.term[0], stage, p.col_offA, lrow0);
      for (int t = 1; t < p.pred_fast.nterms; t++) {
        const unsigned ft = cmp_term<K, FAST && F64ONLY>(p.pred_fast.term[t], stage, p.col_offA, lrow0);
        flags = p.pred_fast.conn[t] ? (flags | ft) : (flags & ft);
      }
      flags &= valid;
      // ranks of this lane's selected rows inside the warp's slice of the tile (row order: k major, lane minor)
      unsigned pos[K];
      unsigned run = 0;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, (flags >> k) & 1u);
        pos[k] = run + __popc(m & lt_mask);
        run += __popc(m);
      }
      unsigned long long* slab_w = p.slab + (((size_t)blockIdx.x * p.slab_slots + (size_t)(it % p.slab_slots)) * p.nproj) * TILE + (size_t)warp * 32 * K;
      for (int q = 0; q < p.nproj; q++) {
        const FastOp& fo = p.proj_fast[q];
        unsigned long long v[K];
        if (fo.kind < 2) {
          const unsigned long long* A = (const unsigned long long*)(stage + p.col_offA[fo.a]) + lrow0;
#pragma unroll
          for (int k = 0; k < K; k++) v[k] = A[k * 32];
        } else if (F64ONLY || fo.ty == DFGPU_FLOAT64) {
          double o[K];
          arith_term_t<K, double>(fo, stage, p.col_offA, lrow0, u2d(fo.imm), flags, bad, o);
#pragma unroll
          for (int k = 0; k < K; k++) v[k] = d2u(o[k]);
        } else {  // Int64 / UInt64: two's complement wrap-around is the natural 64-bit result
          unsigned long long o[K];
          arith_term_t<K, unsigned long long>(fo, stage, p.col_offA, lrow0, fo.imm, flags, bad, o);
#pragma unroll
          for (int k = 0; k < K; k++) v[k] = o[k];
        }
        unsigned long long* dst = slab_w + (size_t)q * TILE;
#pragma unroll
        for (int k = 0; k < K; k++)
          if ((flags >> k) & 1u) dst[pos[k]] = v[k];
      }
      __syncwarp();  // the warp's slab writes are ordered before its later reads (pass 2 runs lanes over other lanes' values)
      if (lane == 0) {
        mbar_arrive(&sh.emptyA[s]);
        sh.s_cnt[it % TM_RING][warp] = run;
        mbar_arrive(&sh.cnt_ready[it % TM_RING]);
      }
      return run;
    };
    // pass 2: the warp's `cnt` stashed values go to their final place; coalesced both ways
    auto stash2 = [&](int it, unsigned cnt) {
      mbar_wait(&sh.pfx_ready[it % TM_RING], (it / TM_RING) & 1);
      const unsigned long long base = sh.s_off[it % TM_RING][warp];
      const unsigned long long* slab_w = p.slab + (((size_t)blockIdx.x * p.slab_slots + (size_t)(it % p.slab_slots)) * p.nproj) * TILE + (size_t)warp * 32 * K;
      for (int q = 0; q < p.nproj; q++) {
        const unsigned long long* src = slab_w + (size_t)q * TILE;
        unsigned long long* dst = (unsigned long long*)p.out[q] + base;
        unsigned long long v[K];
#pragma unroll
        for (int k = 0; k < K; k++)
          if ((unsigned)(k * 32 + lane) < cnt) v[k] = __ldcg(src + k * 32 + lane);
#pragma unroll
        for (int k = 0; k < K; k++)
          if ((unsigned)(k * 32 + lane) < cnt) dst[k * 32 + lane] = v[k];
      }
    };

    int nloc = 0;
    for (int tile = first; tile < p.ntiles; tile += step) nloc++;
    int sb = 0;
    unsigned phb = 0;
    if (STASH) {
      int sa = 0;
      unsigned pha = 0;
      for (int it = 0; it < nloc + LAG; it++) {
        if (it < nloc) {
          stash1(it, first + it * step, sa, pha);
          if (++sa == SA) { sa = 0; pha ^= 1u; }
        }
        if (it >= LAG && !p.noscan) stash2(it - LAG, sh.s_cnt[(it - LAG) % TM_RING][warp]);  // the warp's own count, published LAG tiles ago
      }
    } else if (!p.has_pred) {
      // pure projection: no predicate pass, no lag
      for (int it = 0; it < nloc; it++) {
        const int tile = first + it * step;
        unsigned valid = KMASK;
        if (tile == p.ntiles - 1) {
          const long long row0 = (long long)tile * TILE + warp * 32 * K + lane;
          valid = 0;
#pragma unroll
          for (int k = 0; k < K; k++)
            if (row0 + k * 32 < p.nrows) valid |= 1u << k;
        }
        phase2(it, tile, valid, sb, phb);
        if (++sb == SB) { sb = 0; phb ^= 1u; }
      }
    } else {
      // software pipeline: predicate of tile it, projections of tile it - LAG; the flag bits of the
      // last LAG+1 tiles live in a 128-bit shift register (K bits per tile: up to 15 tiles of lag at K = 8)
      unsigned __int128 fl = 0;
      int sa = 0;
      unsigned pha = 0;
      for (int it = 0; it < nloc + LAG; it++) {
        unsigned f0 = 0;
        if (it < nloc) {
          f0 = phase1(it, first + it * step, sa, pha);
          if (++sa == SA) { sa = 0; pha ^= 1u; }
        }
        fl = (fl << K) | (unsigned __int128)f0;
        if (it >= LAG) {
          phase2(it - LAG, first + (it - LAG) * step, (unsigned)(fl >> (K * LAG)) & KMASK, sb, phb);
          if (++sb == SB) { sb = 0; phb ^= 1u; }
        }
      }
    }
    if (bad) *p.err_flag = 1u;
  }
}

template <int DEPTH, int K, bool F64ONLY, bool FAST, bool STASH = false, int LEAN = 0>
static void launch_one(dfgpu_ctx* ctx, const FPParams& p, size_t smem) {
  auto kern = k_filter_project_tma<DEPTH, K, F64ONLY, FAST, STASH, LEAN>;
  if (ctx->first_use((const void*)kern))
    DF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BUDGET + 16384 + TM_HDR_BYTES));
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
  bool fast = !p.has_pred || p.pred_fast.nterms > 0;
  for (int q = 0; q < p.nproj; q++) fast = fast && p.proj_fast[q].kind > 0;
  bool all_f64 = true;
  for (int c = 0; c < p.ps.ncols; c++) all_f64 = all_f64 && p.ps.cols[c].dtype == DFGPU_FLOAT64;
  // lean shapes: one comparison over 8-byte operands (Float64 / Int64 / UInt64), one or two copy / arithmetic projections over 8-byte columns (DFGPU_FP_LEAN=0: A/B switch)
  auto w8 = [](int dt) { return dt == DFGPU_FLOAT64 || dt == DFGPU_INT64 || dt == DFGPU_UINT64; };
  bool lean = fast && p.has_pred && p.pred_fast.nterms == 1 && w8(p.pred_fast.term[0].ty) && p.nproj >= 1 && p.nproj <= LEAN_MAX_PROJ && !p.slab && !p.count_ballot;
  for (int q = 0; lean && q < p.nproj; q++) lean = w8(p.proj_fast[q].ty);  // copies and arithmetic over 8-byte columns only
  if (const char* e = getenv("DFGPU_FP_LEAN")) lean = lean && atoi(e) != 0;
  if (lean && p.nproj == 1) launch_one<1, K, true, true, false, 1>(ctx, p, smem);
  else if (lean) launch_one<1, K, true, true, false, 2>(ctx, p, smem);
  else if (p.slab && all_f64) launch_one<1, K, true, true, true>(ctx, p, smem);
  else if (p.slab) launch_one<1, K, false, true, true>(ctx, p, smem);
  else if (fast && all_f64) launch_one<1, K, true, true>(ctx, p, smem);
  else if (fast) launch_one<1, K, false, true>(ctx, p, smem);
  else if (p.ps.f64_only) launch_one<DEPTH, K, true, false>(ctx, p, smem);
  else launch_one<DEPTH, K, false, false>(ctx, p, smem);
}

bool launch_fp_tma(dfgpu_ctx* ctx, FPParams& p) {
  if (p.ps.max_depth > 4 || p.ps.ncols < 1) return false;
  // which column slots does the predicate read, which do the projections read?
  bool inA[kMaxCols] = {}, inB[kMaxCols] = {};
  for (int prog = 0; prog < p.ps.nprog; prog++) {
    bool* dst = (p.has_pred && prog == 0) ? inA : inB;
    for (int pc = p.ps.start[prog]; pc < p.ps.start[prog + 1]; pc++) {
      const DevInsn& di = p.ps.insn[pc];
      if (di.op == V_PUSH_COL || (di.op > V_CAST && di.mode == RHS_COL)) dst[di.slot] = true;
    }
  }
  int rowA = 0, rowB = 0;
  for (int c = 0; c < p.ps.ncols; c++) {
    p.col_w[c] = dtype_width(p.ps.cols[c].dtype);
    if (p.col_w[c] <= 0) return false;
    if (inA[c]) rowA += p.col_w[c];
    if (inB[c]) rowB += p.col_w[c];
  }
  // projections of literals only, or a predicate over literals only: leave to the direct kernel
  if (rowB == 0 || (p.has_pred && rowA == 0)) return false;
  // Two layouts.
  //  single ring: one ring holds the UNION of the referenced columns; a tile stays staged from its
  //    predicate pass until its projection pass LAG tiles later (no second load).  Needs LAG + 2 stages,
  //    so it is used when the union is narrow enough for >= 6 stages of a >= 1024-row tile.
  //  dual ring: ring A = predicate columns, ring B = projection columns re-read LAG tiles later from
  //    L2 (evict_last / evict_first hints).  Any width; the lag can be long.
  int rowU = 0;
  for (int c = 0; c < p.ps.ncols; c++)
    if (inA[c] || inB[c]) rowU += p.col_w[c];
  const char* mode = getenv("DFGPU_FP_MODE");  // experiment knob: single | dual | stash
  int K = 0;
  p.single_ring = 0;
  p.slab = nullptr;
  p.slab_slots = 0;
  p.noscan = getenv("DFGPU_FP_NOSCAN") && atoi(getenv("DFGPU_FP_NOSCAN")) != 0;
  // stash mode: every program a fast shape over 8-byte columns with 8-byte results
  // (opt-in, DFGPU_FP_MODE=stash: measured slower than the dual ring at 50 % selectivity, equal at 1 %: profiles/r02m_sweep_fp.txt)
  bool stash = p.has_pred && p.pred_fast.nterms > 0 && mode && std::string(mode) == "stash";
  for (int q = 0; stash && q < p.nproj; q++) {
    const FastOp& fo = p.proj_fast[q];
    stash = fo.kind > 0 && dtype_width(p.ps.out_dtype[q + p.has_pred]) == 8 &&
            (fo.kind < 2 || fo.ty == DFGPU_FLOAT64 || fo.ty == DFGPU_INT64 || fo.ty == DFGPU_UINT64);
  }
  for (int c = 0; stash && c < p.ps.ncols; c++) stash = !(inA[c] || inB[c]) || p.col_w[c] == 8;
  if (stash) {
    for (int k : {8, 4, 2}) {
      const long long tile = (long long)TM_CWARPS * 32 * k;
      if (tile * rowU * 3 <= TM_SMEM_BUDGET) { K = k; break; }
    }
    if (!K) stash = false;
  }
  if (stash) {
    const int tile = TM_CWARPS * 32 * K;
    int off = 0;
    for (int c = 0; c < p.ps.ncols; c++) {
      p.col_offA[c] = (inA[c] || inB[c]) ? off : -1;
      p.col_offB[c] = -1;
      if (inA[c] || inB[c]) off += tile * p.col_w[c];
    }
    p.stage_bytesA = off;
    p.stage_bytesB = 0;
    p.nstagesA = std::min(TM_MAX_STAGES, TM_SMEM_BUDGET / off);
    p.nstagesB = 0;
    const long long grid = std::min(ctx->sm_count, TM_MAX_GRID);
    // lag: the slab slots the grid keeps live (lag x grid x nproj x tile x 8 bytes, about half of it touched at
    // 50 % selectivity) should stay L2 resident
    const long long slot_bytes = grid * p.nproj * (long long)tile * 8;
    p.lag = (int)std::min<long long>(TM_MAX_LAG, std::max<long long>(TM_BATCH, (96ll << 20) / slot_bytes));
    if (const char* e = getenv("DFGPU_FP_LAG")) {
      const int l = atoi(e);
      if (l >= TM_BATCH - 1 && l <= TM_MAX_LAG) p.lag = l;
    }
    p.slab_slots = p.lag + 2;
    p.slab = (unsigned long long*)ctx->alloc(size_t(grid) * size_t(p.slab_slots) * size_t(slot_bytes / grid));
    p.ntiles = int((p.nrows + tile - 1) / tile);
    p.count_ballot = 0;
    const size_t smem = TM_HDR_BYTES + (size_t)p.nstagesA * p.stage_bytesA;
    if (K == 8) launch_k<2, 8>(ctx, p, smem);
    else if (K == 4) launch_k<2, 4>(ctx, p, smem);
    else launch_k<2, 2>(ctx, p, smem);
    ctx->free(p.slab);  // stream ordered: the block is only handed out again to work queued behind this kernel
    return true;
  }
  if (p.has_pred && mode && std::string(mode) == "single") {  // measured slower than dual on B200 (profiles/r01_history.md): opt-in only
    for (int k : {8, 4, 2}) {
      if (k == 8 && p.ps.max_depth > 2) continue;
      const long long tile = (long long)TM_CWARPS * 32 * k;
      if (tile * rowU * 6 <= TM_SMEM_BUDGET + 16384) { K = k; p.single_ring = 1; break; }
    }
  }
  if (!p.single_ring) {
    // rows per lane K in {8,4,2}: the biggest tile that still gives both rings 3 stages.  Measured on
    // B200 (profiles/r01_microbench_fp.txt): per-tile fixed costs (barrier hand-offs, offset gather)
    // outweigh deeper prefetch, so bigger tiles with few stages beat smaller tiles with many.
    for (int k : {8, 4, 2}) {
      if (k == 8 && p.ps.max_depth > 2) continue;  // deep register stacks spill at 8 rows per lane
      const long long tile = (long long)TM_CWARPS * 32 * k;
      if (tile * (rowA + rowB) * 3 <= TM_SMEM_BUDGET) { K = k; break; }
    }
  }
  if (const char* e = getenv("DFGPU_FP_K")) {  // experiment knob
    const int k = atoi(e);
    if (!p.single_ring && (k == 8 || k == 4 || k == 2) && (long long)TM_CWARPS * 32 * k * (rowA + rowB) * 2 <= TM_SMEM_BUDGET && !(k == 8 && p.ps.max_depth > 2)) K = k;
  }
  if (!K) return false;
  const int tile = TM_CWARPS * 32 * K;
  if (p.single_ring)
    for (int c = 0; c < p.ps.ncols; c++) inA[c] = inB[c] = inA[c] || inB[c];
  int offA = 0, offB = 0;
  for (int c = 0; c < p.ps.ncols; c++) {
    // tile is a multiple of 512 rows: every column slice stays 128-B aligned
    p.col_offA[c] = inA[c] ? offA : -1;
    if (inA[c]) offA += tile * p.col_w[c];
    p.col_offB[c] = inB[c] ? offB : -1;
    if (inB[c]) offB += tile * p.col_w[c];
  }
  p.stage_bytesA = offA;
  p.stage_bytesB = offB;
  if (p.single_ring) {
    p.nstagesA = std::min(TM_MAX_STAGES, (TM_SMEM_BUDGET + 16384) / offA);
    p.nstagesB = p.nstagesA;  // the projection pass walks the same ring
    p.stage_bytesB = 0;
    p.lag = p.nstagesA - 3;  // LAG+1 stages are held by the consumers, 2 are prefetch depth
    if (p.lag < TM_BATCH - 1) return false;
    if (const char* e = getenv("DFGPU_FP_LAG")) {
      const int l = atoi(e);
      if (l >= TM_BATCH - 1 && l <= p.nstagesA - 2) p.lag = l;
    }
  } else {
    const int S = std::min(TM_MAX_STAGES, TM_SMEM_BUDGET / (offA + offB));
    p.nstagesA = offA ? S : 0;
    p.nstagesB = S;
    if (!p.has_pred) p.nstagesB = std::min(TM_MAX_STAGES, TM_SMEM_BUDGET / offB);
    if (const char* e = getenv("DFGPU_FP_STAGES")) {  // experiment knob: "A,B"
      int sa = 0, sb = 0;
      if (sscanf(e, "%d,%d", &sa, &sb) == 2 && sa >= 1 && sb >= 1 && sa <= TM_MAX_STAGES && sb <= TM_MAX_STAGES &&
          (long long)sa * offA + (long long)sb * offB <= TM_SMEM_BUDGET + 16384 && p.has_pred) {
        p.nstagesA = sa;
        p.nstagesB = sb;
      }
    }
    // lag: as large as the flag shift register allows (K bits per tile in 64 bits), but the bytes the
    // projection stream will re-read (lag x grid x stage B) must still be in L2 when it gets there
    p.lag = 0;
    if (p.has_pred) {
      const long long l2_budget = 40ll << 20;
      const long long per_tile = (long long)std::min(ctx->sm_count, TM_MAX_GRID) * offB;
      p.lag = (int)std::min<long long>(std::min(TM_MAX_LAG, 128 / K - 1), std::max<long long>(4, l2_budget / per_tile));
    }
    if (const char* e = getenv("DFGPU_FP_LAG")) {  // experiment knob
      const int l = atoi(e);
      if (p.has_pred && l >= TM_BATCH - 1 && l <= std::min(TM_MAX_LAG, 128 / K - 1)) p.lag = l;
    }
  }
  p.ntiles = int((p.nrows + tile - 1) / tile);
  {
    const char* e = getenv("DFGPU_FP_COUNT");
    p.count_ballot = (e && std::string(e) == "ballot") ? 1 : 0;
  }
  const size_t smem = TM_HDR_BYTES + (size_t)p.nstagesA * p.stage_bytesA + (size_t)p.nstagesB * p.stage_bytesB;
  const int d = p.ps.max_depth;
  if (K == 8) launch_k<2, 8>(ctx, p, smem);
  else if (K == 4) { if (d <= 2) launch_k<2, 4>(ctx, p, smem); else launch_k<4, 4>(ctx, p, smem); }
  else { if (d <= 2) launch_k<2, 2>(ctx, p, smem); else launch_k<4, 2>(ctx, p, smem); }
  return true;
}

}  // namespace dfgpu
