// filter_project.cu — FilterRelation + ProjectRelation as ONE order-preserving stream-compaction
// kernel (K1 predicate-eval + K2 filter-gather + K3 fused expr/project of SURVEY.md §2b).
//
// Reference path replaced (per batch): predicate closure -> BooleanArray (src/execution/filter.rs:50),
// per-column builder gather of EVERY input column (filter.rs:55-57,79-110), then one closure pass
// per projection expression (src/execution/projection.rs:49-50).  Here: one pass over the referenced
// columns only; the predicate lives in registers as ballot masks, never in HBM; projected values are
// computed in registers and written straight to their compacted position.
//
// Order preservation (filter.rs:86-90 appends in row order) = single-pass chained scan with
// decoupled look-back across tiles; tiles are claimed through an atomic ticket so that a tile's
// predecessors are always resident (forward progress without any grid-wide barrier).
#include <memory>

#include "filter_project.cuh"

namespace dfgpu {

// NULLS: some referenced column has a validity bitmap.  The predicate is evaluated with arrow's null
// semantics (a null And/Or result reads as false, like `filter.value(i)` in filter.rs:86).  With a
// predicate the projections then see null-free arrays, exactly like the reference: `fn filter` copies
// values and drops the bitmap (filter.rs:83-91) before ProjectRelation runs.  Without a predicate the
// projections run on the original arrays and their validity is written out (32 rows per ballot word).
template <int DEPTH, bool NULLS>
__global__ void __launch_bounds__(FP_THREADS) k_filter_project(const __grid_constant__ FPParams p) {
  __shared__ int s_tile;
  __shared__ unsigned s_wcount[FP_ITEMS * FP_WARPS];
  __shared__ unsigned s_woff[FP_ITEMS * FP_WARPS];
  __shared__ unsigned long long s_prefix;
  static_assert(FP_ITEMS * FP_WARPS == 64, "scan below assumes 64 (item,warp) counters = 2 per lane");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (;;) {
    if (tid == 0) s_tile = (int)atomicAdd(p.ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= p.ntiles) break;
    const long long base = (long long)tile * FP_TILE;

    // ---- phase 1: predicate -> per-thread flag bits (bit j = row base + j*THREADS + tid) ----
    unsigned flags = 0;
    bool bad = false;
#pragma unroll 1
    for (int c = 0; c < FP_CHUNKS; c++) {
      GlobalRows<FP_R> src;
      src.valid = 0;
#pragma unroll
      for (int r = 0; r < FP_R; r++) {
        long long row = base + (long long)(c * FP_R + r) * FP_THREADS + tid;
        src.rows[r] = row < p.nrows ? row : -1;
        if (row < p.nrows) src.valid |= 1u << r;
      }
      if (p.has_pred) {
        unsigned long long v[FP_R];
        unsigned ov;
        unsigned b = eval_program_n<DEPTH, FP_R, false, NULLS>(p.ps, 0, src, v, ov);
        bad = bad || (b != 0);
#pragma unroll
        for (int r = 0; r < FP_R; r++)
          if (((src.valid >> r) & 1u) && (v[r] & 1ull)) flags |= 1u << (c * FP_R + r);
      } else {
        flags |= src.valid << (c * FP_R);
      }
    }
#pragma unroll
    for (int j = 0; j < FP_ITEMS; j++) {
      unsigned b = __ballot_sync(0xffffffffu, (flags >> j) & 1u);
      if (lane == 0) s_wcount[j * FP_WARPS + warp] = __popc(b);
    }
    __syncthreads();

    // ---- warp 0: scan the 64 (item,warp) counts, then chain to the preceding tiles ----
    if (warp == 0) {
      unsigned c0 = s_wcount[2 * lane], c1 = s_wcount[2 * lane + 1];
      unsigned incl = c0 + c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      unsigned excl = incl - (c0 + c1);
      s_woff[2 * lane] = excl;
      s_woff[2 * lane + 1] = excl + c0;
      const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);

      unsigned long long prefix = 0;
      if (!p.has_pred) {
        prefix = (unsigned long long)base;  // nothing is dropped: positions are known without chaining
      } else if (tile == 0) {
        if (lane == 0) st_relaxed(&p.tile_status[0], ST_INCL | total);
      } else {
        if (lane == 0) st_relaxed(&p.tile_status[tile], ST_AGG | total);
        long long look = (long long)tile - 1;
        unsigned long long run = 0;
        for (;;) {
          const long long idx = look - lane;
          unsigned long long s = idx >= 0 ? ld_relaxed(&p.tile_status[idx]) : ST_INCL;  // before tile 0: inclusive 0
          while (__any_sync(0xffffffffu, (s >> 62) == 0)) {
            if ((s >> 62) == 0) s = ld_relaxed(&p.tile_status[idx]);
          }
          const unsigned incl_mask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
          if (incl_mask) {
            const int first = __ffs(incl_mask) - 1;  // nearest predecessor holding an inclusive prefix
            run += warp_sum64(lane <= first ? (s & ST_MASK) : 0ull);
            break;
          }
          run += warp_sum64(s & ST_MASK);
          look -= 32;
        }
        prefix = run;
        if (lane == 0) st_relaxed(&p.tile_status[tile], ST_INCL | (prefix + total));
      }
      if (lane == 0) {
        s_prefix = prefix;
        if (tile == p.ntiles - 1) *p.out_count = prefix + total;
      }
    }
    __syncthreads();

    // ---- phase 2: evaluate the projections, write selected rows at their compacted position ----
    const unsigned long long prefix = s_prefix;
    for (int q = 0; q < p.nproj; q++) {
      const int prog = q + p.has_pred;
      const int odt = p.ps.out_dtype[prog];
      void* o = p.out[q];
#pragma unroll 1
      for (int c = 0; c < FP_CHUNKS; c++) {
        GlobalRows<FP_R> src;
        src.valid = 0;
#pragma unroll
        for (int r = 0; r < FP_R; r++) {
          long long row = base + (long long)(c * FP_R + r) * FP_THREADS + tid;
          src.rows[r] = row < p.nrows ? row : -1;
          if (row < p.nrows) src.valid |= 1u << r;
        }
        unsigned long long v[FP_R];
        unsigned ov = (1u << FP_R) - 1u;
        unsigned b;
        if (NULLS && !p.has_pred) b = eval_program_n<DEPTH, FP_R, false, true>(p.ps, prog, src, v, ov);
        else b = eval_program<DEPTH, FP_R, false>(p.ps, prog, src, v);
        if (NULLS && !p.has_pred && p.out_valid[q]) {
#pragma unroll
          for (int r = 0; r < FP_R; r++) {
            const bool inb = (src.valid >> r) & 1u;
            const unsigned vb = __ballot_sync(0xffffffffu, inb && ((ov >> r) & 1u));
            const unsigned ib = __ballot_sync(0xffffffffu, inb);
            if (lane == 0 && ib) {
              p.out_valid[q][src.rows[r] >> 5] = vb;  // the warp's 32 rows are consecutive and 32-aligned
              const unsigned nn = __popc(ib & ~vb);
              if (nn) atomicAdd(&p.null_counts[q], (unsigned long long)nn);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < FP_R; r++) {
          const int j = c * FP_R + r;
          const bool f = (flags >> j) & 1u;
          const unsigned m = __ballot_sync(0xffffffffu, f);
          if (f) {
            const unsigned long long idx = prefix + s_woff[j * FP_WARPS + warp] + __popc(m & lt_mask);
            store_elem(o, odt, (long long)idx, v[r]);
            // DivideByZero only counts for rows that survive the filter: ProjectRelation runs on
            // the filtered batch (src/execution/context.rs:140-161).
            if ((b >> r) & 1u) bad = true;
          }
        }
      }
    }
    if (bad) *p.err_flag = 1u;
    __syncthreads();  // s_tile / s_wcount are reused by the next tile
  }
}

// one byte per row -> BooleanArray bits (LSB first); one warp packs 32 rows into one word
__global__ void __launch_bounds__(256) k_pack_bits(const unsigned char* __restrict__ bytes, long long n, unsigned* __restrict__ words) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long padded = (n + 31) / 32 * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += stride) {
    const unsigned m = __ballot_sync(0xffffffffu, i < n && bytes[i] != 0);
    if ((threadIdx.x & 31) == 0) words[i >> 5] = m;
  }
}
static int grid_for_rows(dfgpu_ctx* ctx, long long rows) {
  long long g = (rows + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 8;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

template <int DEPTH, bool NULLS = false>
static void launch_fp(dfgpu_ctx* ctx, const FPParams& p) {
  int per_sm = 0;
  DF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_filter_project<DEPTH, NULLS>, FP_THREADS, 0));
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > p.ntiles) grid = p.ntiles;
  const int ps = ctx->prof_begin();
  k_filter_project<DEPTH, NULLS><<<(unsigned)grid, FP_THREADS, 0, ctx->stream>>>(p);
  DF_CUDA(cudaGetLastError());
  ctx->prof_end(ps);
  ctx->launches++;
}

}  // namespace dfgpu

namespace dfgpu {
void gather_utf8(dfgpu_ctx* ctx, const DevColumn& src, const unsigned long long* d_idx, long long nsel, DevColumn* out);
}
using namespace dfgpu;

// Host-only type check of one expression program (no ctx, no device): the same ProgramBuilder the
// operators use, over a schema-only stand-in for a batch.
extern "C" int dfgpu_check_program(const int32_t* col_dtypes, int ncols, const dfgpu_insn* prog, int prog_len, int32_t* out_dtype) {
  return guarded([&] {
    if (!col_dtypes || ncols < 0 || !prog || prog_len <= 0 || !out_dtype) fail(DFGPU_ERR_GENERAL, "dfgpu_check_program: null argument");
    dfgpu_batch schema_only;  // ctx == nullptr: owns nothing, its destructor frees nothing
    for (int i = 0; i < ncols; i++) {
      DevColumn c;
      c.dtype = col_dtypes[i];
      schema_only.cols.push_back(c);
    }
    ProgramBuilder pb(&schema_only);
    const int pi = pb.add(prog, prog_len, "expression");
    ProgramSet ps;
    pb.finish(&ps);  // instruction / column-slot limits
    *out_dtype = pb.out_dtype(pi);
  });
}

extern "C" int dfgpu_filter_project(dfgpu_ctx* ctx, const dfgpu_batch* batch, const dfgpu_insn* pred, int pred_len,
                                    const dfgpu_insn* const* proj, const int* proj_len, int nproj, dfgpu_result** out) {
  return guarded([&] {
    if (!ctx || !batch || !out) fail(DFGPU_ERR_GENERAL, "dfgpu_filter_project: null argument");
    ctx->use();
    ProgramBuilder pb(batch);
    const int has_pred = pred_len > 0 ? 1 : 0;
    if (has_pred) {
      int pi = pb.add(pred, pred_len, "predicate");
      if (pb.out_dtype(pi) != DFGPU_BOOL)  // filter.rs:64-66
        fail(DFGPU_ERR_EXECUTION, "Filter expression did not evaluate to boolean");
    }
    // nproj == 0: FilterRelation alone emits every input column (filter.rs:55-57)
    std::vector<dfgpu_insn> ident;
    std::vector<const dfgpu_insn*> pptr;
    std::vector<int> plen;
    if (nproj == 0) {
      ident.resize(batch->cols.size());
      for (size_t i = 0; i < batch->cols.size(); i++) {
        memset(&ident[i], 0, sizeof(dfgpu_insn));
        ident[i].op = DFGPU_OP_COL;
        ident[i].col = int(i);
        ident[i].dtype = batch->cols[i].dtype;
      }
      for (size_t i = 0; i < batch->cols.size(); i++) {
        pptr.push_back(&ident[i]);
        plen.push_back(1);
      }
      nproj = int(batch->cols.size());
      proj = pptr.data();
      proj_len = plen.data();
    }
    // Projections that are a plain Utf8 column are gathered by row number after the fused kernel
    // (utf8_gather.cu); everything else is evaluated inside it.
    std::vector<int> out_kind;  // per output column: >= 0 kernel program slot, -1 - c = Utf8 gather of input column c
    int nkern = 0;
    bool any_utf8 = false;
    for (int i = 0; i < nproj; i++) {
      if (proj_len[i] == 1 && proj[i][0].op == DFGPU_OP_COL && proj[i][0].col >= 0 && size_t(proj[i][0].col) < batch->cols.size() &&
          batch->cols[size_t(proj[i][0].col)].dtype == DFGPU_UTF8) {
        out_kind.push_back(-1 - proj[i][0].col);
        any_utf8 = true;
        continue;
      }
      int pi = pb.add(proj[i], proj_len[i], "projection");
      int dt = pb.out_dtype(pi);
      if (!is_numeric(dt) && dt != DFGPU_BOOL)
        fail(DFGPU_ERR_NOT_IMPLEMENTED, std::string("filter/projection output of type ") + dtype_name(dt) +
                                            " is not supported on the GPU path yet");
      out_kind.push_back(nkern++);
    }
    int rowid_slot = -1;
    if (any_utf8) {
      pb.add_rowid();
      rowid_slot = nkern++;
    }
    // columns referenced anywhere must be null-free and fixed width for now
    FPParams p;
    pb.finish(&p.ps);
    for (int s = 0; s < p.ps.ncols; s++) {
      // Boolean input columns (bit-packed BooleanArray) are read by the direct kernel's interpreter
      if (!is_numeric(p.ps.cols[s].dtype) && p.ps.cols[s].dtype != DFGPU_BOOL)
        fail(DFGPU_ERR_NOT_IMPLEMENTED, std::string("expressions over ") + dtype_name(p.ps.cols[s].dtype) + " columns are not supported on the GPU path yet");
    }
    if (p.ps.max_depth > 8) fail(DFGPU_ERR_NOT_IMPLEMENTED, "expression too deep (register stack depth > 8)");
    // Boolean projections (comparisons / AND / OR, expression.rs:212-224,236-290) leave the kernel as one
    // byte per selected row and are bit-packed (BooleanArray layout) once the row count is known.
    for (int k = 0; k < nkern; k++)
      if (p.ps.out_dtype[k + has_pred] == DFGPU_BOOL) p.ps.out_dtype[k + has_pred] = DFGPU_UINT8;

    auto res = std::make_unique<dfgpu_result>();
    res->ctx = ctx;
    const long long n = batch->nrows;
    dfgpu_result bool_bytes;  // RAII for the unpacked Boolean outputs
    bool_bytes.ctx = ctx;
    std::vector<int> bool_of_out(size_t(nproj), -1);
    // kernel outputs (worst case n rows each); the row-number column is scratch, not a result column
    dfgpu_result scratch;  // RAII for the row-number buffer
    scratch.ctx = ctx;
    std::vector<void*> kern_out(size_t(nkern), nullptr);
    for (int i = 0; i < nproj; i++) {
      DevColumn c;
      if (out_kind[size_t(i)] >= 0) {
        c.dtype = pb.out_dtype(out_kind[size_t(i)] + has_pred);
        if (c.dtype == DFGPU_BOOL) {
          DevColumn b;
          b.dtype = DFGPU_UINT8;
          b.values_bytes = size_t(n > 0 ? n : 1);
          b.values = ctx->alloc(b.values_bytes);
          bool_of_out[size_t(i)] = int(bool_bytes.cols.size());
          bool_bytes.cols.push_back(b);
          kern_out[size_t(out_kind[size_t(i)])] = b.values;
          c.values_bytes = size_t((n + 31) / 32) * 4 + 4;  // packed, whole 32-bit words
          c.values = ctx->alloc(c.values_bytes);
        } else {
          c.values_bytes = size_t(n > 0 ? n : 1) * size_t(dtype_width(c.dtype));
          c.values = ctx->alloc(c.values_bytes);
          kern_out[size_t(out_kind[size_t(i)])] = c.values;
        }
      } else {
        c.dtype = DFGPU_UTF8;  // filled by the gather below
      }
      res->cols.push_back(c);
    }
    if (rowid_slot >= 0) {
      DevColumn c;
      c.dtype = DFGPU_UINT64;
      c.values_bytes = size_t(n > 0 ? n : 1) * 8;
      c.values = ctx->alloc(c.values_bytes);
      scratch.cols.push_back(c);
      kern_out[size_t(rowid_slot)] = c.values;
    }
    if (n == 0) {
      for (auto& c : res->cols)
        if (c.dtype == DFGPU_BOOL) c.values_bytes = 0;
      for (auto& c : res->cols)
        if (c.dtype == DFGPU_UTF8) {
          c.offsets = (int32_t*)ctx->alloc(4);
          DF_CUDA(cudaMemsetAsync(c.offsets, 0, 4, ctx->stream));
          c.values = ctx->alloc(1);
        }
      DF_CUDA(cudaStreamSynchronize(ctx->stream));
      res->nrows = 0;
      *out = res.release();
      return;
    }
    p.nrows = n;
    p.has_pred = has_pred;
    p.nproj = nkern;
    for (int i = 0; i < nkern; i++) p.out[i] = kern_out[size_t(i)];
    // tile_status is sized for the smallest tile either kernel uses (1024 rows)
    const size_t max_tiles = size_t((n + 1023) / 1024) + 1;
    // one allocation, one memset: [ticket, pad..] then the tile words.  The row count and the error
    // flag are written by the kernel straight into pinned host memory (zero-copy), so the step ends
    // with a stream synchronise and no device-to-host copy.
    unsigned long long* status = (unsigned long long*)ctx->alloc((max_tiles + 8) * 8);
    DF_CUDA(cudaMemsetAsync(status, 0, (max_tiles + 8) * 8, ctx->stream));
    ctx->h_scratch[0] = 0;
    ctx->h_scratch[2] = 0;
    p.tile_status = status + 8;
    p.ticket = (unsigned*)(status + 0);
    p.out_count = ctx->h_scratch + 0;
    p.err_flag = (unsigned*)(ctx->h_scratch + 2);
    // validity outputs: only a query WITHOUT a predicate can emit nulls (see k_filter_project)
    memset(p.out_valid, 0, sizeof(p.out_valid));
    p.null_counts = ctx->d_scratch + 32;
    std::vector<int> valid_of_out(size_t(nproj), -1);  // result column -> kernel program with a validity buffer
    if (p.ps.has_nulls && !has_pred) {
      DF_CUDA(cudaMemsetAsync(ctx->d_scratch + 32, 0, kMaxProgs * 8, ctx->stream));
      for (int i = 0; i < nproj; i++) {
        const int k = out_kind[size_t(i)];
        if (k >= 0 && p.ps.nullable[k]) {
          const size_t vbytes = size_t((n + 31) / 32) * 4;
          res->cols[size_t(i)].validity = (uint8_t*)ctx->alloc(vbytes);
          DF_CUDA(cudaMemsetAsync(res->cols[size_t(i)].validity, 0, vbytes, ctx->stream));
          p.out_valid[k] = (unsigned*)res->cols[size_t(i)].validity;
          valid_of_out[size_t(i)] = k;
        }
      }
    }
    // fast shapes (see FPParams): straight from the lowered bytecode
    auto fast_type = [&](int dt) {
      return dt == DFGPU_FLOAT64 || dt == DFGPU_INT64 || dt == DFGPU_UINT64 || dt == DFGPU_FLOAT32 || dt == DFGPU_INT32 || dt == DFGPU_UINT32;
    };
    auto fast_of = [&](int prog, bool is_pred) {
      FastOp f;
      memset(&f, 0, sizeof(f));
      const int b = p.ps.start[prog], e = p.ps.start[prog + 1];
      const DevInsn* in = &p.ps.insn[b];
      if (in[0].op != V_PUSH_COL) return f;
      const int dt = p.ps.cols[in[0].slot].dtype;
      if (e - b == 1 && !is_pred) {  // plain column copy: any fixed width
        f.kind = 1;
        f.a = in[0].slot;
        f.ty = dt;
        return f;
      }
      if (e - b != 2 || in[1].mode == RHS_STACK || !fast_type(dt)) return f;
      const bool cmp = in[1].op >= V_EQ && in[1].op <= V_GE, arith = in[1].op >= V_ADD && in[1].op <= V_DIV;
      if (is_pred ? !cmp : !arith) return f;
      if (arith && (dt == DFGPU_INT32 || dt == DFGPU_UINT32)) return f;               // narrow wrap-around: interpreter
      if (arith && in[1].op == V_DIV && !(dt == DFGPU_FLOAT64 || dt == DFGPU_FLOAT32)) return f;  // integer division: interpreter
      if (in[1].mode == RHS_COL && p.ps.cols[in[1].slot].dtype != dt) return f;
      f.kind = in[1].mode == RHS_COL ? 2 : 3;
      f.op = in[1].op;
      f.a = in[0].slot;
      f.b = in[1].slot;
      f.ty = dt;
      f.imm = in[1].imm;
      return f;
    };
    memset(&p.pred_fast, 0, sizeof(p.pred_fast));
    if (has_pred) {
      // t0 [t1 AND|OR [t2 AND|OR ...]] in lowered form: (PUSH_COL, CMP leaf) {(PUSH_COL, CMP leaf), AND|OR stack}*
      const int b = p.ps.start[0], e = p.ps.start[1];
      const DevInsn* in = &p.ps.insn[b];
      auto term_at = [&](int i, FastOp* out) {
        if (i + 1 >= e - b) return false;
        const DevInsn &c = in[i], &o = in[i + 1];
        if (c.op != V_PUSH_COL || !fast_type(p.ps.cols[c.slot].dtype)) return false;
        if (o.op < V_EQ || o.op > V_GE || o.mode == RHS_STACK) return false;
        if (o.mode == RHS_COL && p.ps.cols[o.slot].dtype != p.ps.cols[c.slot].dtype) return false;
        memset(out, 0, sizeof(*out));
        out->kind = o.mode == RHS_COL ? 2 : 3;
        out->op = o.op;
        out->a = c.slot;
        out->b = o.slot;
        out->ty = p.ps.cols[c.slot].dtype;
        out->imm = o.imm;
        return true;
      };
      FastPred fp;
      memset(&fp, 0, sizeof(fp));
      int i = 0;
      bool ok = term_at(0, &fp.term[0]);
      fp.nterms = ok ? 1 : 0;
      i = 2;
      while (ok && i < e - b) {
        if (fp.nterms >= 4 || !term_at(i, &fp.term[fp.nterms]) || i + 2 >= e - b) { ok = false; break; }
        const DevInsn& j = in[i + 2];
        if ((j.op != V_AND && j.op != V_OR) || j.mode != RHS_STACK) { ok = false; break; }
        fp.conn[fp.nterms] = j.op == V_OR ? 1 : 0;
        fp.nterms++;
        i += 3;
      }
      if (ok) p.pred_fast = fp;
    }
    for (int i = 0; i < nkern; i++) p.proj_fast[i] = fast_of(i + has_pred, false);
    if (p.ps.has_nulls) {
      p.ntiles = int((n + FP_TILE - 1) / FP_TILE);
      launch_fp<8, true>(ctx, p);  // the null-aware evaluator lives in the direct kernel only
    } else if (ctx->force_direct_kernel || !launch_fp_tma(ctx, p)) {
      p.ntiles = int((n + FP_TILE - 1) / FP_TILE);
      const int d = p.ps.max_depth;
      if (d <= 1) launch_fp<1>(ctx, p);
      else if (d <= 2) launch_fp<2>(ctx, p);
      else if (d <= 4) launch_fp<4>(ctx, p);
      else launch_fp<8>(ctx, p);
    }
    if (p.ps.has_nulls && !has_pred)
      DF_CUDA(cudaMemcpyAsync(ctx->h_scratch + 32, ctx->d_scratch + 32, kMaxProgs * 8, cudaMemcpyDeviceToHost, ctx->stream));
    DF_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->free(status);
    for (int i = 0; i < nproj; i++) {
      DevColumn& c = res->cols[size_t(i)];
      if (valid_of_out[size_t(i)] >= 0) {
        c.null_count = (int64_t)ctx->h_scratch[32 + valid_of_out[size_t(i)]];
        if (c.null_count == 0) {
          ctx->free(c.validity);
          c.validity = nullptr;
        }
      }
    }
    if ((unsigned)ctx->h_scratch[2] != 0) fail(DFGPU_ERR_ARROW, "DivideByZero");
    res->nrows = has_pred ? (int64_t)ctx->h_scratch[0] : n;
    for (int i = 0; i < nproj; i++)
      if (bool_of_out[size_t(i)] >= 0) {
        DevColumn& c = res->cols[size_t(i)];
        const long long words = (res->nrows + 31) / 32;
        if (words > 0) {
          k_pack_bits<<<grid_for_rows(ctx, words * 32), 256, 0, ctx->stream>>>(
              (const unsigned char*)bool_bytes.cols[size_t(bool_of_out[size_t(i)])].values, res->nrows, (unsigned*)c.values);
          DF_CUDA(cudaGetLastError());
          ctx->launches++;
        }
        c.values_bytes = size_t(res->nrows + 7) / 8;
        any_utf8 = true;  // synchronise before the byte buffers are released
      }
    for (int i = 0; i < nproj; i++)
      if (out_kind[size_t(i)] < 0)
        gather_utf8(ctx, batch->cols[size_t(-1 - out_kind[size_t(i)])], (const unsigned long long*)scratch.cols[0].values, res->nrows,
                    &res->cols[size_t(i)]);
    if (!has_pred)
      for (int i = 0; i < nproj; i++)
        if (out_kind[size_t(i)] < 0) {  // Utf8 column passed through untouched keeps its validity (expression.rs:313)
          const DevColumn& srcc = batch->cols[size_t(-1 - out_kind[size_t(i)])];
          if (srcc.null_count > 0) {
            const size_t vb = size_t(n + 7) / 8;
            res->cols[size_t(i)].validity = (uint8_t*)ctx->alloc(vb);
            DF_CUDA(cudaMemcpyAsync(res->cols[size_t(i)].validity, srcc.validity, vb, cudaMemcpyDeviceToDevice, ctx->stream));
            res->cols[size_t(i)].null_count = srcc.null_count;
          }
        }
    if (any_utf8) DF_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = res.release();
  });
}

// ---------------------------------------------------------------------------------------------
// Host -> host, chunk-pipelined (see include/dfgpu.h).  All H2D copies are queued up front on the
// copy-in stream; chunk c's kernel waits for its copies only; its compacted output is copied back
// on the copy-out stream while chunk c+1 is being filtered and chunk c+2 uploaded.
// ---------------------------------------------------------------------------------------------
extern "C" int dfgpu_filter_project_host(dfgpu_ctx* ctx, const dfgpu_col* cols, int ncols, const dfgpu_insn* pred, int pred_len,
                                         const dfgpu_insn* const* proj, const int* proj_len, int nproj, int64_t chunk_rows,
                                         dfgpu_result** out) {
  return guarded([&] {
    if (!ctx || !out || !cols || ncols < 1) fail(DFGPU_ERR_GENERAL, "dfgpu_filter_project_host: null argument");
    ctx->use();
    const long long n = cols[0].len;
    for (int i = 0; i < ncols; i++) {
      if (cols[i].len != n) fail(DFGPU_ERR_GENERAL, "all columns of a RecordBatch must have the same length");
    }
    // nproj == 0 -> every input column
    std::vector<dfgpu_insn> ident;
    std::vector<const dfgpu_insn*> pptr;
    std::vector<int> plen;
    if (nproj == 0) {
      ident.resize(size_t(ncols));
      for (int i = 0; i < ncols; i++) {
        memset(&ident[size_t(i)], 0, sizeof(dfgpu_insn));
        ident[size_t(i)].op = DFGPU_OP_COL;
        ident[size_t(i)].col = i;
        ident[size_t(i)].dtype = cols[i].dtype;
        pptr.push_back(&ident[size_t(i)]);
        plen.push_back(1);
      }
      nproj = ncols;
      proj = pptr.data();
      proj_len = plen.data();
    }
    // referenced columns only (the reference uploads nothing, but gathers every column: filter.rs:55-57)
    std::vector<int> remap(size_t(ncols), -1), used;
    auto scan = [&](const dfgpu_insn* p, int len) {
      for (int i = 0; i < len; i++)
        if (p[i].op == DFGPU_OP_COL) {
          if (p[i].col < 0 || p[i].col >= ncols) fail(DFGPU_ERR_INVALID_COLUMN, "column index " + std::to_string(p[i].col) + " out of range");
          if (remap[size_t(p[i].col)] < 0) {
            remap[size_t(p[i].col)] = int(used.size());
            used.push_back(p[i].col);
          }
        }
    };
    if (pred_len > 0) scan(pred, pred_len);
    for (int q = 0; q < nproj; q++) scan(proj[q], proj_len[q]);
    if (used.empty()) fail(DFGPU_ERR_NOT_IMPLEMENTED, "queries that reference no column");
    // Utf8 / Boolean / nullable inputs: not chunk-pipelined — the referenced columns are uploaded whole and the resident
    // operator runs (same kernels, same results); the result then lives in DEVICE memory (dfgpu_result_on_host says
    // which; dfgpu_result_copy_col works for both)
    bool resident = false;
    for (int c : used) resident = resident || !is_numeric(cols[c].dtype) || cols[c].validity;
    auto rewrite = [&](const dfgpu_insn* p, int len) {
      std::vector<dfgpu_insn> v(p, p + len);
      for (auto& in : v)
        if (in.op == DFGPU_OP_COL) in.col = remap[size_t(in.col)];
      return v;
    };
    std::vector<dfgpu_insn> pred2 = pred_len > 0 ? rewrite(pred, pred_len) : std::vector<dfgpu_insn>();
    std::vector<std::vector<dfgpu_insn>> proj2;
    std::vector<const dfgpu_insn*> proj2p;
    std::vector<int> proj2l;
    for (int q = 0; q < nproj; q++) {
      proj2.push_back(rewrite(proj[q], proj_len[q]));
    }
    for (auto& v : proj2) {
      proj2p.push_back(v.data());
      proj2l.push_back(int(v.size()));
    }
    auto run_resident = [&] {
      std::vector<dfgpu_col> sub;
      for (int c : used) sub.push_back(cols[c]);
      dfgpu_batch* b = nullptr;
      int rc = dfgpu_batch_upload(ctx, sub.data(), int(sub.size()), &b);
      if (rc != DFGPU_OK) fail(rc, dfgpu_last_error());
      struct G { dfgpu_batch* b; ~G() { dfgpu_batch_free(b); } } g{b};
      rc = dfgpu_filter_project(ctx, b, pred2.data(), int(pred2.size()), proj2p.data(), proj2l.data(), nproj, out);
      if (rc != DFGPU_OK) fail(rc, dfgpu_last_error());
    };
    if (resident) {
      run_resident();
      return;
    }

    // default chunk: measured on C2 (profiles/r02y_e2e_chunks.txt): 16 / 8 / 4 / 2 / 1 Mi rows -> 16.59 / 15.96 / 15.45 / 15.47 / 15.99 ms per step
    if (chunk_rows <= 0) chunk_rows = 4ll << 20;
    const long long nchunks = n > 0 ? (n + chunk_rows - 1) / chunk_rows : 1;
    struct Chunk {
      dfgpu_batch batch;
      dfgpu_result* res = nullptr;
      cudaEvent_t ev = nullptr;
    };
    std::vector<std::unique_ptr<Chunk>> chunks;
    struct Cleanup {
      std::vector<std::unique_ptr<Chunk>>* c;
      dfgpu_ctx* ctx;
      ~Cleanup() {
        cudaStreamSynchronize(ctx->stream_in);
        cudaStreamSynchronize(ctx->stream_out);
        for (auto& ch : *c) {
          if (ch->res) delete ch->res;
          if (ch->ev) cudaEventDestroy(ch->ev);
        }
      }
    } cleanup{&chunks, ctx};
    // make the allocations (ordered on ctx->stream) visible to the copy-in stream
    cudaEvent_t ev_alloc;
    DF_CUDA(cudaEventCreateWithFlags(&ev_alloc, cudaEventDisableTiming));
    for (long long c = 0; c < nchunks; c++) {
      auto ch = std::make_unique<Chunk>();
      ch->batch.ctx = ctx;
      const long long r0 = c * (long long)chunk_rows, rows = std::min<long long>((long long)chunk_rows, n - r0);
      ch->batch.nrows = rows > 0 ? rows : 0;
      for (int u : used) {
        DevColumn d;
        d.dtype = cols[u].dtype;
        d.values_bytes = size_t(ch->batch.nrows) * size_t(dtype_width(d.dtype));
        d.values = ctx->alloc(d.values_bytes);
        ch->batch.cols.push_back(d);
      }
      DF_CUDA(cudaEventCreateWithFlags(&ch->ev, cudaEventDisableTiming));
      chunks.push_back(std::move(ch));
    }
    DF_CUDA(cudaEventRecord(ev_alloc, ctx->stream));
    DF_CUDA(cudaStreamWaitEvent(ctx->stream_in, ev_alloc, 0));
    cudaEventDestroy(ev_alloc);
    for (long long c = 0; c < nchunks; c++) {
      Chunk& ch = *chunks[size_t(c)];
      const long long r0 = c * chunk_rows;
      for (size_t k = 0; k < used.size(); k++) {
        const dfgpu_col& hc = cols[used[k]];
        const int w = dtype_width(hc.dtype);
        if (ch.batch.nrows > 0)
          DF_CUDA(cudaMemcpyAsync(ch.batch.cols[k].values, static_cast<const uint8_t*>(hc.values) + size_t(hc.offset + r0) * size_t(w),
                                  size_t(ch.batch.nrows) * size_t(w), cudaMemcpyHostToDevice, ctx->stream_in));
      }
      DF_CUDA(cudaEventRecord(ch.ev, ctx->stream_in));
    }
    // filter chunk by chunk; outputs go back as soon as their size is known
    auto res = std::make_unique<dfgpu_result>();
    res->ctx = ctx;
    res->on_host = true;
    long long off = 0;
    for (long long c = 0; c < nchunks; c++) {
      Chunk& ch = *chunks[size_t(c)];
      DF_CUDA(cudaStreamWaitEvent(ctx->stream, ch.ev, 0));
      int rc = dfgpu_filter_project(ctx, &ch.batch, pred2.data(), int(pred2.size()), proj2p.data(), proj2l.data(), nproj, &ch.res);
      if (rc != 0) fail(rc, dfgpu_last_error());
      if (c == 0) {
        for (int q = 0; q < nproj; q++) {
          DevColumn hcol;
          hcol.dtype = ch.res->cols[size_t(q)].dtype;
          if (!is_numeric(hcol.dtype)) {  // Boolean projections (bit-packed): the resident operator handles the whole batch
            resident = true;
            break;
          }
          hcol.values_bytes = size_t(n > 0 ? n : 1) * size_t(dtype_width(hcol.dtype));
          hcol.values = ctx->host_alloc(hcol.values_bytes);
          res->cols.push_back(hcol);
        }
      }
      if (resident) break;
      // the kernel of this chunk has completed (dfgpu_filter_project synchronised ctx->stream)
      for (int q = 0; q < nproj; q++) {
        const int w = dtype_width(res->cols[size_t(q)].dtype);
        if (ch.res->nrows > 0)
          DF_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(res->cols[size_t(q)].values) + size_t(off) * size_t(w), ch.res->cols[size_t(q)].values,
                                  size_t(ch.res->nrows) * size_t(w), cudaMemcpyDeviceToHost, ctx->stream_out));
      }
      off += ch.res->nrows;
    }
    DF_CUDA(cudaStreamSynchronize(ctx->stream_out));
    if (resident) {
      res.reset();  // returns the pinned blocks of the projections before the Boolean one
      run_resident();
      return;
    }
    res->nrows = off;
    *out = res.release();
  });
}
