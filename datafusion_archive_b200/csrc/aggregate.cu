// aggregate.cu — AggregateRelation on the GPU (K4 column reduce, K5 hash-aggregate, K7 table
// compaction, K6 partial-aggregate merge of SURVEY.md §2b).
//
// Reference path replaced: `with_group_by` (src/execution/aggregate.rs:787-952) — per row a
// heap-allocated Vec<GroupByScalar> key, an FNV hash-map lookup, and a boxed ScalarValue folded
// through Rc<RefCell<dyn AggregateFunction>> (aggregate.rs:548-612, 102-283) — and
// `without_group_by` (aggregate.rs:703-785).  Here the map is an open-addressed table in HBM
// (linear probing, 64-bit packed keys claimed with atomicCAS) whose accumulators are updated with
// fire-and-forget L2 reductions: RED.ADD.F64 for f64 SUM, RED.ADD.U64 for COUNT / integer SUM,
// RED.MIN/MAX.U64 on an order-preserving encoding for MIN/MAX (bit-exact for every type).
#include <memory>

#include "expr_vm.cuh"

namespace dfgpu {

void utf8_hash(dfgpu_ctx* ctx, const DevColumn& src, long long n, unsigned long long* d_out);
void shift_copy_i32(dfgpu_ctx* ctx, int* dst, const int* src, long long n, int add);
void gather_utf8_multi(dfgpu_ctx* ctx, const Utf8Source* d_srcs, const unsigned long long* d_idx, long long nsel, DevColumn* out);

constexpr int AG_THREADS = 256;
constexpr int AG_R = 2;  // rows per thread per tile: 2 measured best on B200 (1.65 vs 1.74 ms at 4, 1.94 at 8: the kernel is bound by scattered L2 reductions, not by loads in flight)
constexpr int AG_TILE = AG_THREADS * AG_R;
constexpr int kMaxAggs = 8;
constexpr int kMaxKeys = 4;
constexpr unsigned long long EMPTY_KEY = ~0ull;
constexpr int AG_MAX_PROBE = 1 << 14;
constexpr long long AG_MIN_CAP = 1ll << 22;

struct AggDesc {
  uint8_t func;   // DFGPU_AGG_*
  uint8_t mtype;  // machine type of the argument
  uint8_t dtype;  // Arrow dtype of the argument
  uint8_t out_dtype;
};

// Table addressing.  A slot is a LINE of lw 64-bit words (lw a power of two; word 0 = the packed key) plus
// one word in each of n_add separate arrays.  loc[a] says where accumulator a lives: >= 1 = that word of
// the line, < 0 = additive array ~loc[a].  Three layouts fall out of it (measured on B200,
// profiles/r02a_scatter_ops2.txt and profiles/r01m_microbench_agg.txt):
//   SoA     lw = 1, every accumulator in its own array.  A row's probe and reductions go to different L2
//           slices in parallel; reductions that hit the SAME 32-byte sector as the probe serialise in the
//           slice (LDG + 2 RED on one sector: 2.04 ms per 1e8 rows against 1.38 ms on three arrays).
//   hybrid  MIN / MAX accumulators share the line with the key, SUM / COUNT stay in arrays.  The probe is
//           one 128/256-bit load that also returns the current MIN / MAX, and a MIN / MAX reduction is only
//           issued when the row improves on the value just read (monotone accumulators: a stale read can
//           only cause a redundant reduction, never a missed one).  After a group's first few rows almost
//           no row does, so MIN + MAX + SUM costs one load and one reduction per row instead of one load
//           and three reductions.
//   line    every accumulator in the line (lw >= 1 + naggs): one sector per group; wins once the table no
//           longer fits L2 and every touched sector is an HBM transaction (1e7 groups: 7.1 vs 11.2 ms).
struct TableLayout {
  unsigned long long* base;  // (cap + 1) lines of lw words
  unsigned long long* add;   // n_add arrays of (cap + 1) words
  long long lw, astride;
  signed char loc[kMaxAggs];
  __host__ __device__ __forceinline__ unsigned long long* key(long long slot) const { return base + slot * lw; }
  __host__ __device__ __forceinline__ unsigned long long* val(long long slot, int a) const {
    const int l = loc[a];
    return l >= 0 ? base + slot * lw + l : add + (long long)(~l) * astride + slot;
  }
};

// plain-column fast path (k_hash_agg_plain): every key and aggregate argument is a plain column and the
// WHERE clause, if any, is a chain of column comparisons: no interpreter in the kernel
struct PlainTerm {
  int a, b;      // column slots (b: right-hand column when kind == 2)
  int kind;      // 2 = col cmp col, 3 = col cmp imm
  int op;        // VOp (V_EQ .. V_GE)
  int mt;        // machine type of the operands
  int conn;      // joins the running result with this term: 0 = AND, 1 = OR
  unsigned long long imm;
};
struct PlainSpec {
  int ncols;                  // distinct column slots referenced (<= 4): ps.cols[0 .. ncols)
  int key_slot[kMaxKeys];
  int arg_slot[kMaxAggs];     // per distinct argument program
  int nterms;                 // predicate terms (0 = no predicate)
  PlainTerm term[4];
};

struct AggParams {
  ProgramSet ps;  // programs [0,nkeys) = group keys, then the distinct aggregate-argument programs
  AggDesc aggs[kMaxAggs];
  // aggregates over the same argument expression (MIN(v), MAX(v), SUM(v)) share one evaluation:
  // programs [nkeys, nkeys + nargs) are the DISTINCT argument programs, agg_arg[a] picks one
  int agg_arg[kMaxAggs];
  int nargs;
  unsigned long long key_mask[kMaxKeys];
  int key_shift[kMaxKeys];
  int nkeys, naggs;
  long long nrows;
  long long row_begin;       // process rows [row_begin, row_begin + nrows) of the batch
  const unsigned* row_list;  // non-null: process rows row_list[0..nlist) (overflow replay)
  long long nlist;
  // cap+1 slots; slot cap is reserved for the key that equals EMPTY_KEY
  TableLayout t;
  long long cap;             // power of two
  long long max_groups;      // new keys are refused (-> overflow list) beyond this fill
  unsigned long long* counters;  // [0] ngroups [1] overflow count [2] sentinel-key used [3] error
  unsigned* ovf_rows;
  // front table (FRONT kernels): one table of front_slots per CTA, or — for very few groups — one
  // private table per warp, so that shared-memory atomics only contend inside a warp
  int front_slots;     // power of two
  int front_per_warp;  // 0 / 1
  int stream_hint;     // 1: input columns are loaded with an L2 evict-first policy
  // fused WHERE (Aggregate{input: Selection}, context.rs:126-139,162-192): program 0 is the predicate and
  // the key / argument programs follow; rows that fail it are skipped before the probe
  int has_pred;
  int cond_mm;         // 1: MIN / MAX reductions are skipped when the value read with the probe already covers the row
  int table_hint;      // 1: table loads / reductions carry an L2 evict-last policy
  // multi-pass scan (tables that outgrow L2): launch `pass_id` of `npass` takes the rows whose hash's top
  // log2(npass) bits equal pass_id, i.e. whose home slot lies in one contiguous 1/npass of the table
  int npass, pass_id, pass_shift;
  // wide keys (k_hash_agg_wide): composite keys of more than 64 bits and keys with Utf8 parts.  Line word 0 is a
  // TAG (the 64-bit hash of the key tuple, low bit = ready), words 1..kw the key parts: the 64-bit value of a
  // fixed-width part, or for a Utf8 part a reference (source << 40 | row) to the string of the row that created
  // the group.  Equal tags are confirmed by comparing every part (strings byte by byte), so hash collisions
  // just probe on — the reference's Vec<GroupByScalar> equality (aggregate.rs:65-76, 807-852).
  struct {
    int kw;
    int is_utf8[kMaxKeys];
    const int* off[kMaxKeys];            // this batch's Utf8 key columns
    const unsigned char* bytes[kMaxKeys];
    unsigned long long ref_base[kMaxKeys];  // (source index of this batch's column) << UTF8_SRC_SHIFT
    const Utf8Source* srcs;              // every retained Utf8 key column
  } wide;
  // lean kernel (k_hash_agg_lean): one 8-byte key column, one 8-byte argument column, no predicate
  struct {
    const unsigned long long* key_col;
    const unsigned long long* arg_col;
    unsigned long long* sum_arr;  // additive arrays of SUM / COUNT
    unsigned long long* cnt_arr;
    int min_w, max_w;             // words of MIN / MAX in the line
  } lean;
  PlainSpec plain;
};

// ---- order-preserving encodings so that MIN/MAX are native u64 atomics -----------------------
__device__ __forceinline__ unsigned long long ord_enc(unsigned long long v, int mt) {
  switch (mt) {
    case MT_F64: return (v >> 63) ? ~v : (v ^ 0x8000000000000000ull);
    case MT_F32: { unsigned b = (unsigned)v; return (b >> 31) ? (unsigned long long)(~b) : (unsigned long long)(b ^ 0x80000000u); }
    case MT_I: return v ^ 0x8000000000000000ull;
    default: return v;
  }
}
__host__ __device__ __forceinline__ unsigned long long ord_dec(unsigned long long e, int mt) {
  switch (mt) {
    case MT_F64: return (e >> 63) ? (e ^ 0x8000000000000000ull) : ~e;
    case MT_F32: { unsigned b = (unsigned)e; return (b >> 31) ? (unsigned long long)(b ^ 0x80000000u) : (unsigned long long)(~b); }
    case MT_I: return e ^ 0x8000000000000000ull;
    default: return e;
  }
}
__device__ __forceinline__ bool is_nan_val(unsigned long long v, int mt) {
  if (mt == MT_F64) { double d = u2d(v); return d != d; }
  if (mt == MT_F32) { float f = u2f(v); return f != f; }
  return false;
}
__host__ __device__ __forceinline__ unsigned long long agg_identity(int func) {
  return func == DFGPU_AGG_MIN ? ~0ull : 0ull;
}

// fold one value into an accumulator held in a register / shared memory
__device__ __forceinline__ unsigned long long acc_fold(int func, int mt, unsigned long long acc, unsigned long long v) {
  switch (func) {
    case DFGPU_AGG_SUM:
      if (mt == MT_F64) return d2u(u2d(acc) + u2d(v));
      if (mt == MT_F32) return f2u(u2f(acc) + u2f(v));
      return acc + v;
    case DFGPU_AGG_COUNT: return acc + 1ull;
    case DFGPU_AGG_MIN: {
      if (is_nan_val(v, mt)) return acc;  // f64::min ignores NaN (aggregate.rs:139-140)
      unsigned long long e = ord_enc(v, mt);
      return e < acc ? e : acc;
    }
    default: {
      if (is_nan_val(v, mt)) return acc;
      unsigned long long e = ord_enc(v, mt);
      return e > acc ? e : acc;
    }
  }
}
// combine two accumulators
__device__ __forceinline__ unsigned long long acc_merge(int func, int mt, unsigned long long a, unsigned long long b) {
  switch (func) {
    case DFGPU_AGG_SUM:
      if (mt == MT_F64) return d2u(u2d(a) + u2d(b));
      if (mt == MT_F32) return f2u(u2f(a) + u2f(b));
      return a + b;
    case DFGPU_AGG_COUNT: return a + b;
    case DFGPU_AGG_MIN: return a < b ? a : b;
    default: return a > b ? a : b;
  }
}
// combine an accumulator into global memory (fire-and-forget reductions)
__device__ __forceinline__ void acc_merge_global(int func, int mt, unsigned long long* p, unsigned long long b) {
  switch (func) {
    case DFGPU_AGG_SUM:
      if (mt == MT_F64) atomicAdd((double*)p, u2d(b));
      else if (mt == MT_F32) atomicAdd((float*)p, u2f(b));
      else atomicAdd(p, b);
      break;
    case DFGPU_AGG_COUNT: atomicAdd(p, b); break;
    case DFGPU_AGG_MIN: atomicMin(p, b); break;
    default: atomicMax(p, b); break;
  }
}
// fold one raw value into global memory
__device__ __forceinline__ void acc_fold_global(int func, int mt, unsigned long long* p, unsigned long long v) {
  switch (func) {
    case DFGPU_AGG_SUM:
      if (mt == MT_F64) atomicAdd((double*)p, u2d(v));
      else if (mt == MT_F32) atomicAdd((float*)p, u2f(v));
      else atomicAdd(p, v);
      break;
    case DFGPU_AGG_COUNT: atomicAdd(p, 1ull); break;
    case DFGPU_AGG_MIN:
      if (!is_nan_val(v, mt)) atomicMin(p, ord_enc(v, mt));
      break;
    default:
      if (!is_nan_val(v, mt)) atomicMax(p, ord_enc(v, mt));
      break;
  }
}

// Home slot of a key = the TOP log2(cap) bits of its hash.  The top bits of the slot index are then the top
// bits of the hash whatever the capacity, which is what the multi-pass scan selects rows by (a pass touches
// one contiguous 1/P of the table, and the assignment of rows to passes survives a table growth).
__host__ __device__ __forceinline__ int hash_shift(long long cap) {
  int lg = 0;
  while ((1ll << lg) < cap) lg++;
  return 64 - lg;
}
__device__ __forceinline__ unsigned long long home_slot(unsigned long long hash, int hshift) { return hshift >= 64 ? 0ull : hash >> hshift; }
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// One table line as read by a probe: word 0 = key, words 1..3 = the accumulators that share the line.
// `pol` != 0: an L2 evict-last policy word — the table competes for L2 with a 1.6 GB input stream that is
// read once (marked evict-first), so its lines should be the last to go.
struct Line {
  unsigned long long w[4];
};
template <bool WITH_VALS>
__device__ __forceinline__ void load_line(const TableLayout& t, long long slot, Line& ln, unsigned long long pol = 0ull) {
  const unsigned long long* q = t.key(slot);
  if (WITH_VALS && t.lw >= 4) {  // 256-bit load (LDG.E.ENL2.256): lines are 32-byte aligned
    if (pol)
      asm volatile("ld.global.cg.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;" : "=l"(ln.w[0]), "=l"(ln.w[1]), "=l"(ln.w[2]), "=l"(ln.w[3]) : "l"(q), "l"(pol) : "memory");
    else
      asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(ln.w[0]), "=l"(ln.w[1]), "=l"(ln.w[2]), "=l"(ln.w[3]) : "l"(q) : "memory");
  } else if (WITH_VALS && t.lw == 2) {
    if (pol)
      asm volatile("ld.global.cg.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(ln.w[0]), "=l"(ln.w[1]) : "l"(q), "l"(pol) : "memory");
    else
      asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(ln.w[0]), "=l"(ln.w[1]) : "l"(q) : "memory");
    ln.w[2] = ln.w[3] = 0ull;
  } else {
    if (pol)
      asm volatile("ld.global.cg.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(ln.w[0]) : "l"(q), "l"(pol) : "memory");
    else
      asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(ln.w[0]) : "l"(q) : "memory");
    ln.w[1] = ln.w[2] = ln.w[3] = 0ull;
  }
}
// fire-and-forget reductions with an L2 cache-policy operand
__device__ __forceinline__ void red_add_f64(unsigned long long* p, double v, unsigned long long pol) {
  asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_f32(unsigned long long* p, float v, unsigned long long pol) {
  asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v, unsigned long long pol) {
  asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_min_u64(unsigned long long* p, unsigned long long v, unsigned long long pol) {
  asm volatile("red.global.min.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_max_u64(unsigned long long* p, unsigned long long v, unsigned long long pol) {
  asm volatile("red.global.max.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned long long line_word(const Line& ln, int l) {
  return l == 1 ? ln.w[1] : (l == 2 ? ln.w[2] : ln.w[3]);
}

// Find the slot of `key`, claiming an empty one when the key is new.  `ln` is the line read at the home
// slot h on entry and the line of the returned slot on exit (as it was when this thread read it: a slot
// claimed meanwhile still shows its initial accumulator identities, which is what the conditional
// MIN / MAX update needs).  Returns -1 when the key is new and the table refuses new keys (fill limit /
// probe limit): the row goes to the overflow list.
template <bool WITH_VALS>
__device__ __forceinline__ long long probe_insert(const TableLayout& t, long long cap, unsigned long long key, Line& ln,
                                                  unsigned long long h, bool full, unsigned& new_groups, unsigned long long pol = 0ull) {
  const unsigned long long mask = (unsigned long long)cap - 1ull;
  for (int probes = 0; probes < AG_MAX_PROBE; ++probes) {
    if (ln.w[0] == key) return (long long)h;
    if (ln.w[0] == EMPTY_KEY) {
      if (full) return -1;
      const unsigned long long old = atomicCAS(t.key((long long)h), EMPTY_KEY, key);
      if (old == EMPTY_KEY) { new_groups++; return (long long)h; }
      if (old == key) return (long long)h;
    }
    h = (h + 1ull) & mask;
    load_line<WITH_VALS>(t, (long long)h, ln, pol);
  }
  return -1;
}

// fold one raw value into global memory; `cur` = the accumulator as read with the probe (have_cur): a
// MIN / MAX that the row does not improve needs no reduction (the stored value only moves towards it)
__device__ __forceinline__ void acc_fold_global_cond(int func, int mt, unsigned long long* p, unsigned long long v, bool have_cur,
                                                     unsigned long long cur, unsigned long long pol) {
  if (func == DFGPU_AGG_MIN) {
    if (is_nan_val(v, mt)) return;
    const unsigned long long e = ord_enc(v, mt);
    if (!have_cur || e < cur) { if (pol) red_min_u64(p, e, pol); else atomicMin(p, e); }
  } else if (func == DFGPU_AGG_MAX) {
    if (is_nan_val(v, mt)) return;
    const unsigned long long e = ord_enc(v, mt);
    if (!have_cur || e > cur) { if (pol) red_max_u64(p, e, pol); else atomicMax(p, e); }
  } else if (!pol) {
    acc_fold_global(func, mt, p, v);
  } else if (func == DFGPU_AGG_COUNT) {
    red_add_u64(p, 1ull, pol);
  } else if (mt == MT_F64) {
    red_add_f64(p, u2d(v), pol);
  } else if (mt == MT_F32) {
    red_add_f32(p, u2f(v), pol);
  } else {
    red_add_u64(p, v, pol);
  }
}

// fold one raw value into a shared-memory accumulator (front table)
__device__ __forceinline__ void acc_fold_shared(int func, int mt, unsigned long long* p, unsigned long long v) {
  switch (func) {
    case DFGPU_AGG_SUM:
      if (mt == MT_F64) atomicAdd((double*)p, u2d(v));
      else if (mt == MT_F32) atomicAdd((float*)p, u2f(v));
      else atomicAdd(p, v);
      break;
    case DFGPU_AGG_COUNT: atomicAdd(p, 1ull); break;
    case DFGPU_AGG_MIN:
      if (!is_nan_val(v, mt)) atomicMin(p, ord_enc(v, mt));
      break;
    default:
      if (!is_nan_val(v, mt)) atomicMax(p, ord_enc(v, mt));
      break;
  }
}

constexpr int AG_FRONT_SLOTS = 2048;   // per-CTA shared-memory front table (low-cardinality GROUP BY)
constexpr int AG_FRONT_PROBES = 8;
constexpr int AG_FRONT_MAX_GROUPS = 1024;

// ---- row sources of the scan kernel -----------------------------------------------------------------
// InterpSrc: keys, arguments and the WHERE predicate are expression programs run by the interpreter of
// expr_vm.cuh, AG_R rows per thread.
// NULLS: like the reference, keys and MIN/MAX/SUM arguments are read ignoring the validity bitmap
// (`array.value(row)`, aggregate.rs:561-601, 807-852; a null produced by arithmetic reads as the
// builder's default 0); only COUNT (an extension, §DESIGN) honours nulls.  A predicate that evaluates to
// null reads as its value false (filter.rs:86 `filter.value(i)`).
template <int DEPTH, bool NULLS>
struct InterpSrc {
  static constexpr int R = AG_R;
  static constexpr bool PREFETCH = false;
  GlobalRows<R> g;
  unsigned mask;  // rows to aggregate: in range and passing the predicate
  unsigned bad = 0;
  __device__ __forceinline__ void load(const AggParams& p, long long tb, long long n, int tid, unsigned long long policy) {
    g.valid = 0;
    g.l2_policy = policy;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const long long i = tb + (long long)r * AG_THREADS + tid;
      g.rows[r] = i < n ? (p.row_list ? (long long)p.row_list[i] : p.row_begin + i) : -1;
      if (i < n) g.valid |= 1u << r;
    }
    mask = g.valid;
    bad = 0;
  }
  __device__ __forceinline__ void prepare(const AggParams& p) {
    if (p.has_pred) {
      unsigned long long v[R];
      unsigned pv;
      bad |= eval_program_n<DEPTH, R, false, NULLS>(p.ps, 0, g, v, pv);  // FilterRelation evaluates the predicate on every row
      unsigned keep = 0;
#pragma unroll
      for (int r = 0; r < R; r++) keep |= (unsigned)(v[r] & 1ull) << r;
      mask &= keep;
    }
  }
  __device__ __forceinline__ void key(const AggParams& p, int k, unsigned long long (&v)[R]) {
    unsigned kv;
    bad |= eval_program_n<DEPTH, R, false, NULLS>(p.ps, p.has_pred + k, g, v, kv) & mask;
  }
  // returns the DivideByZero bits of the rows; av = validity bits of the argument values
  __device__ __forceinline__ unsigned arg(const AggParams& p, int gi, unsigned long long (&v)[R], unsigned& av) {
    return eval_program_n<DEPTH, R, false, NULLS>(p.ps, p.has_pred + p.nkeys + gi, g, v, av);
  }
  __device__ __forceinline__ unsigned rowid(int r) const { return (unsigned)g.rows[r]; }
};

// PlainSrc: every key / argument is a plain 4- or 8-byte column and the predicate a chain of column
// comparisons.  A thread owns two CONSECUTIVE rows, so an 8-byte column is one 128-bit load per thread
// (coalesced 512 bytes per warp instruction) and a 4-byte column one 64-bit load; values are kept in the
// widened 64-bit machine representation of the interpreter, so everything downstream is shared.
__device__ __forceinline__ void ld_pair64(const void* base, long long row, bool both, unsigned long long policy, unsigned long long (&out)[2]) {
  const unsigned long long* q = (const unsigned long long*)base + row;
  if (both) {
    if (policy)
      asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(out[0]), "=l"(out[1]) : "l"(q), "l"(policy));
    else
      asm volatile("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(out[0]), "=l"(out[1]) : "l"(q));
  } else {
    out[0] = __ldg(q);
    out[1] = 0ull;
  }
}
__device__ __forceinline__ void ld_pair32(const void* base, long long row, bool both, bool sign, unsigned long long (&out)[2]) {
  const unsigned* q = (const unsigned*)base + row;
  unsigned lo, hi = 0;
  if (both) {
    const uint2 t = __ldg((const uint2*)q);
    lo = t.x;
    hi = t.y;
  } else {
    lo = __ldg(q);
  }
  out[0] = sign ? (unsigned long long)(long long)(int)lo : (unsigned long long)lo;
  out[1] = sign ? (unsigned long long)(long long)(int)hi : (unsigned long long)hi;
}
__device__ __forceinline__ unsigned cmp_bits2(int op, int mt, const unsigned long long (&a)[2], const unsigned long long (&b)[2]) {
  unsigned f = 0;
#define DF_C2(EXPR) _Pragma("unroll") for (int r = 0; r < 2; r++) { const unsigned long long x = a[r], y = b[r]; f |= (unsigned)(EXPR) << r; }
#define DF_C2_OPS(CAST)                           \
  switch (op) {                                   \
    case V_EQ: DF_C2(CAST(x) == CAST(y)) break;   \
    case V_NE: DF_C2(CAST(x) != CAST(y)) break;   \
    case V_LT: DF_C2(CAST(x) < CAST(y)) break;    \
    case V_LE: DF_C2(CAST(x) <= CAST(y)) break;   \
    case V_GT: DF_C2(CAST(x) > CAST(y)) break;    \
    default: DF_C2(CAST(x) >= CAST(y)) break;     \
  }
  switch (mt) {
    case MT_F64: DF_C2_OPS(u2d) break;
    case MT_F32: DF_C2_OPS(u2f) break;
    case MT_I: DF_C2_OPS((long long)) break;
    default: DF_C2_OPS((unsigned long long)) break;
  }
#undef DF_C2_OPS
#undef DF_C2
  return f;
}
// NC: column slots the instantiation holds (2 = the common key + value shape: fewer registers, more CTAs
// per SM); PF: the loads of the NEXT tile are issued before the current tile's probes, so the HBM latency of
// the input stream and the L2 latency of the probe chain overlap instead of adding up per thread.
template <int NC, bool PF>
struct PlainSrc {
  static constexpr int R = 2;
  static constexpr bool PREFETCH = PF;
  unsigned long long cv[NC][2];
  long long row0;
  unsigned mask;
  unsigned bad = 0;
  // value selects instead of indexed access: the column values stay in registers
  __device__ __forceinline__ void col(int s, unsigned long long (&v)[2]) const {
    if (NC == 2) {
      v[0] = s == 0 ? cv[0][0] : cv[1][0];
      v[1] = s == 0 ? cv[0][1] : cv[1][1];
    } else {
      v[0] = s == 0 ? cv[0][0] : (s == 1 ? cv[1][0] : (s == 2 ? cv[2 % NC][0] : cv[3 % NC][0]));
      v[1] = s == 0 ? cv[0][1] : (s == 1 ? cv[1][1] : (s == 2 ? cv[2 % NC][1] : cv[3 % NC][1]));
    }
  }
  __device__ __forceinline__ void load(const AggParams& p, long long tb, long long n, int tid, unsigned long long policy) {
    const long long i = tb + 2ll * tid;
    row0 = p.row_begin + i;  // even: tb and row_begin are even (host-checked)
    mask = (i < n ? 1u : 0u) | (i + 1 < n ? 2u : 0u);
    bad = 0;
    const bool both = mask == 3u;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      cv[c][0] = cv[c][1] = 0ull;
      if (c < p.plain.ncols && mask) {
        const int dt = p.ps.cols[c].dtype;  // warp-uniform
        if (dt == DFGPU_FLOAT64 || dt == DFGPU_INT64 || dt == DFGPU_UINT64) ld_pair64(p.ps.cols[c].ptr, row0, both, policy, cv[c]);
        else ld_pair32(p.ps.cols[c].ptr, row0, both, dt == DFGPU_INT32, cv[c]);
      }
    }
  }
  __device__ __forceinline__ void prepare(const AggParams& p) {
    if (p.plain.nterms > 0) {
      unsigned keep = 0;
      for (int t = 0; t < p.plain.nterms; t++) {
        const PlainTerm& pt = p.plain.term[t];
        unsigned long long x[2], y[2];
        col(pt.a, x);
        if (pt.kind == 2) col(pt.b, y);
        else y[0] = y[1] = pt.imm;
        const unsigned f = cmp_bits2(pt.op, pt.mt, x, y);
        keep = t == 0 ? f : (pt.conn ? (keep | f) : (keep & f));
      }
      mask &= keep;
    }
  }
  __device__ __forceinline__ void key(const AggParams& p, int k, unsigned long long (&v)[2]) { col(p.plain.key_slot[k], v); }
  __device__ __forceinline__ unsigned arg(const AggParams& p, int gi, unsigned long long (&v)[2], unsigned& av) {
    col(p.plain.arg_slot[gi], v);
    av = 3u;
    return 0u;
  }
  __device__ __forceinline__ unsigned rowid(int r) const { return (unsigned)(row0 + r); }
};

// K5 scan.  Per row: key -> mix64 -> linear probing over table lines (ld.global.cg, 64/128/256 bits: the
// probe also brings the in-line MIN / MAX accumulators) -> atomicCAS to claim an empty slot -> one
// fire-and-forget L2 reduction per additive accumulator and per MIN / MAX that the row improves.
// FRONT: a per-CTA open-addressed table in shared memory absorbs the updates (shared-memory atomics),
// and is merged into the global table once, when the CTA is done.  Used when the sampled prefix shows
// few groups: with a handful of hot keys every global reduction would serialise on the same L2 sector
// (measured: 10 groups, 1e8 rows: 21 ms through L2 atomics).
template <class Src, bool FRONT, bool NULLS>
__device__ __forceinline__ void hash_agg_body(const AggParams& p, unsigned long long* s_front) {
  constexpr int R = Src::R;
  const int FS = p.front_slots;                                            // slots per front table
  const int ftables = p.front_per_warp ? AG_THREADS / 32 : 1;              // tables per CTA
  unsigned long long* ftab = s_front + (p.front_per_warp ? (size_t)(threadIdx.x >> 5) * FS * (1 + p.naggs) : 0);
  if (FRONT) {
    for (int i = threadIdx.x; i < FS * ftables; i += AG_THREADS) {
      unsigned long long* tb = s_front + (size_t)(i / FS) * FS * (1 + p.naggs);
      tb[i % FS] = EMPTY_KEY;
      for (int a = 0; a < p.naggs; a++) tb[(1 + a) * FS + (i % FS)] = agg_identity(p.aggs[a].func);
    }
    __syncthreads();
  }
  const int tid = threadIdx.x, lane = tid & 31;
  const long long n = p.row_list ? p.nlist : p.nrows;
  const int hshift = hash_shift(p.cap);
  // the input is read exactly once: mark its lines evict-first so that they do not displace the table
  const unsigned long long stream_policy = p.stream_hint ? l2_evict_first_policy() : 0ull;
  const unsigned long long tpol = p.table_hint ? l2_evict_last_policy() : 0ull;
  bool bad = false;
  constexpr int TILE = AG_THREADS * R;
  const long long tstep = (long long)gridDim.x * TILE;
  Src src, nxt;
  bool first = true;
  for (long long tb = (long long)blockIdx.x * TILE; tb < n; tb += tstep) {
    // fill limit, once per warp per tile (no CTA-wide barrier in the steady state)
    unsigned long long filled = 0;
    if (lane == 0) filled = __ldcg(&p.counters[0]);
    filled = __shfl_sync(0xffffffffu, filled, 0);
    const bool full = (long long)filled >= p.max_groups;
    if (!Src::PREFETCH || first) src.load(p, tb, n, tid, stream_policy);
    first = false;
    if (Src::PREFETCH && tb + tstep < n) nxt.load(p, tb + tstep, n, tid, stream_policy);
    src.prepare(p);
    // group key: one packed 64-bit word (GroupByScalar vector of aggregate.rs:807-852)
    unsigned long long key[R];
#pragma unroll
    for (int r = 0; r < R; r++) key[r] = 0;
    for (int k = 0; k < p.nkeys; k++) {
      unsigned long long v[R];
      src.key(p, k, v);
#pragma unroll
      for (int r = 0; r < R; r++) key[r] |= (v[r] & p.key_mask[k]) << p.key_shift[k];
    }
    // front table (shared memory): claim / find the key there first
    int fslot[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      fslot[r] = -1;
      if (FRONT && ((src.mask >> r) & 1u) && key[r] != EMPTY_KEY) {
        unsigned fh = (unsigned)(mix64(key[r]) >> 40) & (unsigned)(FS - 1);
        for (int pr = 0; pr < AG_FRONT_PROBES; pr++) {
          unsigned long long c = ftab[fh];
          if (c == EMPTY_KEY) c = atomicCAS(&ftab[fh], EMPTY_KEY, key[r]);
          if (c == key[r] || c == EMPTY_KEY) { fslot[r] = (int)fh; break; }
          fh = (fh + 1) & (unsigned)(FS - 1);
        }
      }
    }
    // first probe of all R rows issued back to back (R independent L2 requests in flight)
    unsigned long long h[R];
    Line ln[R];
    bool probing[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const unsigned long long hs = mix64(key[r]);
      h[r] = home_slot(hs, hshift);
      // multi-pass scan: this launch only takes the rows whose home slot lies in its 1/npass of the table
      if (p.npass > 1 && (key[r] == EMPTY_KEY ? p.pass_id != 0 : (int)(hs >> p.pass_shift) != p.pass_id)) src.mask &= ~(1u << r);
      probing[r] = ((src.mask >> r) & 1u) && key[r] != EMPTY_KEY && fslot[r] < 0;
      ln[r].w[0] = ln[r].w[1] = ln[r].w[2] = ln[r].w[3] = 0ull;
      if (probing[r]) load_line<true>(p.t, (long long)h[r], ln[r], tpol);
    }
    long long slot[R];
    unsigned new_groups = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      slot[r] = -1;
      if (!((src.mask >> r) & 1u) || fslot[r] >= 0) continue;
      if (key[r] == EMPTY_KEY) {  // the one key value that collides with the empty marker
        if (__ldcg(&p.counters[2]) == 0ull) p.counters[2] = 1ull;
        slot[r] = p.cap;
        continue;
      }
      slot[r] = probe_insert<true>(p.t, p.cap, key[r], ln[r], h[r], full, new_groups, tpol);
      if (slot[r] < 0) {
        const unsigned long long at = atomicAdd(&p.counters[1], 1ull);
        p.ovf_rows[at] = src.rowid(r);
      }
    }
    // accumulators (update_accumulators, aggregate.rs:548-612): argument evaluated once per row
    for (int g = 0; g < p.nargs; g++) {
      unsigned long long v[R];
      unsigned av;
      const unsigned b = src.arg(p, g, v, av);
      for (int a = 0; a < p.naggs; a++) {
        if (p.agg_arg[a] != g) continue;
        const int func = p.aggs[a].func, mt = p.aggs[a].mtype, l = p.t.loc[a];
        const bool cond = p.cond_mm && l >= 1 && l <= 3 && (func == DFGPU_AGG_MIN || func == DFGPU_AGG_MAX);
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (NULLS && func == DFGPU_AGG_COUNT && !((av >> r) & 1u)) continue;  // COUNT counts non-null values
          if (FRONT && fslot[r] >= 0 && ((src.mask >> r) & 1u)) {
            acc_fold_shared(func, mt, &ftab[(1 + a) * FS + fslot[r]], v[r]);
            if ((b >> r) & 1u) bad = true;
          } else if (slot[r] >= 0) {
            acc_fold_global_cond(func, mt, p.t.val(slot[r], a), v[r], cond && probing[r], cond ? line_word(ln[r], l) : 0ull, tpol);
            if ((b >> r) & 1u) bad = true;
          }
        }
      }
    }
    bad = bad || src.bad != 0;
    // one counter update per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) new_groups += __shfl_xor_sync(0xffffffffu, new_groups, o);
    if (lane == 0 && new_groups) atomicAdd(&p.counters[0], (unsigned long long)new_groups);
    if (Src::PREFETCH) src = nxt;
  }
  if (FRONT) {
    // merge this CTA's front table into the global table.  New keys are always admitted here; the host
    // lowers max_groups of a FRONT launch by grid x front slots, so the table stays at most half full.
    __syncthreads();
    unsigned new_groups = 0;
    for (int i = threadIdx.x; i < FS * ftables; i += AG_THREADS) {
      const unsigned long long* tb = s_front + (size_t)(i / FS) * FS * (1 + p.naggs);
      const int j = i % FS;
      const unsigned long long key = tb[j];
      if (key == EMPTY_KEY) continue;
      const unsigned long long h = home_slot(mix64(key), hshift);
      Line ln;
      load_line<false>(p.t, (long long)h, ln);
      const long long slot = probe_insert<false>(p.t, p.cap, key, ln, h, false, new_groups);
      if (slot < 0) { p.counters[3] = 2ull; continue; }
      for (int a = 0; a < p.naggs; a++)
        acc_merge_global(p.aggs[a].func, p.aggs[a].mtype, p.t.val(slot, a), tb[(1 + a) * FS + j]);
    }
    if (new_groups) atomicAdd(&p.counters[0], (unsigned long long)new_groups);
  }
  if (bad) p.counters[3] = 1ull;
}

template <int DEPTH, bool FRONT, bool NULLS>
__global__ void __launch_bounds__(AG_THREADS) k_hash_agg(const __grid_constant__ AggParams p) {
  extern __shared__ unsigned long long s_front[];  // FRONT: keys[AG_FRONT_SLOTS] then vals[naggs][AG_FRONT_SLOTS]
  hash_agg_body<InterpSrc<DEPTH, NULLS>, FRONT, NULLS>(p, s_front);
}
template <int NC, bool PF, bool FRONT>
__global__ void __launch_bounds__(AG_THREADS, 4) k_hash_agg_plain(const __grid_constant__ AggParams p) {
  extern __shared__ unsigned long long s_front[];
  hash_agg_body<PlainSrc<NC, PF>, FRONT, false>(p, s_front);
}

// K5, lean form: the canonical GROUP BY shape — one 8-byte integer key column, one 8-byte argument column, any
// subset M of {MIN = 1, MAX = 2, SUM = 4, COUNT = 8} over it, no WHERE — with everything that is dynamic in the
// generic body (loops over keys / arguments / aggregates, dtype and function switches, layout arithmetic)
// resolved at compile time: ~60 instructions per row instead of ~300 (ncu: the generic kernels are issue-bound
// well before the LSU / L2 ceiling of their scattered operations).  Same table, same protocol.
template <int M, int MT>
__global__ void __launch_bounds__(AG_THREADS, 5) k_hash_agg_lean(const __grid_constant__ AggParams p) {
  constexpr bool HAS_MIN = (M & 1) != 0, HAS_MAX = (M & 2) != 0, HAS_SUM = (M & 4) != 0, HAS_CNT = (M & 8) != 0;
  constexpr int LW = (HAS_MIN && HAS_MAX) ? 4 : ((HAS_MIN || HAS_MAX) ? 2 : 1);  // the hybrid layout's line for this M (host-checked)
  const int tid = threadIdx.x, lane = tid & 31;
  const long long n = p.nrows;
  const unsigned long long* __restrict__ kc = p.lean.key_col + p.row_begin;
  const unsigned long long* __restrict__ vc = p.lean.arg_col + p.row_begin;
  unsigned long long* const base = p.t.base;
  const int hshift = hash_shift(p.cap);
  const unsigned long long smask = (unsigned long long)p.cap - 1ull;
  const unsigned long long policy = p.stream_hint ? l2_evict_first_policy() : 0ull;
  constexpr int TILE = AG_THREADS * 2;
  const long long tstep = (long long)gridDim.x * TILE;
  // software pipeline: the two 128-bit loads of the NEXT tile are in flight while this tile's probes and
  // reductions are issued (HBM latency of the stream and L2 latency of the probe chain overlap per thread)
  // (measured: a gain for SUM / COUNT shapes, a loss where the 256-bit line of MIN / MAX needs the registers)
  constexpr bool PF = LW == 1;
  unsigned long long nk[2] = {0ull, 0ull}, nv[2] = {0ull, 0ull};
  if (PF) {
    const long long i0 = (long long)blockIdx.x * TILE + 2ll * tid;
    if (i0 < n) {
      ld_pair64(kc, i0, i0 + 1 < n, policy, nk);
      ld_pair64(vc, i0, i0 + 1 < n, policy, nv);
    }
  }
  for (long long tb = (long long)blockIdx.x * TILE; tb < n; tb += tstep) {
    unsigned long long filled = 0;
    if (lane == 0) filled = __ldcg(&p.counters[0]);
    filled = __shfl_sync(0xffffffffu, filled, 0);
    const bool full = (long long)filled >= p.max_groups;
    const long long i = tb + 2ll * tid;
    const bool any = i < n, both = i + 1 < n;
    if (!PF && any) {
      ld_pair64(kc, i, both, policy, nk);
      ld_pair64(vc, i, both, policy, nv);
    }
    unsigned long long k[2] = {nk[0], nk[1]}, v[2] = {nv[0], nv[1]};
    if (PF) {
      const long long j = i + tstep;
      if (j < n) {
        ld_pair64(kc, j, j + 1 < n, policy, nk);
        ld_pair64(vc, j, j + 1 < n, policy, nv);
      }
    }
    unsigned long long slot[2];
    Line ln[2];
    bool act[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const unsigned long long hs = mix64(k[r]);
      slot[r] = home_slot(hs, hshift);
      act[r] = (r == 0 ? any : both) && (p.npass == 1 || (k[r] == EMPTY_KEY ? p.pass_id == 0 : (int)(hs >> p.pass_shift) == p.pass_id));
      if (act[r] && k[r] != EMPTY_KEY) {
        const unsigned long long* q = base + slot[r] * LW;
        if (LW == 4) asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(ln[r].w[0]), "=l"(ln[r].w[1]), "=l"(ln[r].w[2]), "=l"(ln[r].w[3]) : "l"(q) : "memory");
        else if (LW == 2) asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(ln[r].w[0]), "=l"(ln[r].w[1]) : "l"(q) : "memory");
        else asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(ln[r].w[0]) : "l"(q) : "memory");
      }
    }
    unsigned new_groups = 0;
#pragma unroll
    for (int r = 0; r < 2; r++) {
      if (!act[r]) continue;
      bool have_line = true;
      if (k[r] == EMPTY_KEY) {  // the one key value that collides with the empty marker
        if (__ldcg(&p.counters[2]) == 0ull) p.counters[2] = 1ull;
        slot[r] = (unsigned long long)p.cap;
        have_line = false;
      } else {
        bool found = false;
        for (int probes = 0; probes < AG_MAX_PROBE; ++probes) {
          if (ln[r].w[0] == k[r]) { found = true; break; }
          if (ln[r].w[0] == EMPTY_KEY) {
            if (full) break;
            const unsigned long long old = atomicCAS(base + slot[r] * LW, EMPTY_KEY, k[r]);
            if (old == EMPTY_KEY) { new_groups++; found = true; break; }
            if (old == k[r]) { found = true; break; }
          }
          slot[r] = (slot[r] + 1ull) & smask;
          const unsigned long long* q = base + slot[r] * LW;
          if (LW == 4) asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(ln[r].w[0]), "=l"(ln[r].w[1]), "=l"(ln[r].w[2]), "=l"(ln[r].w[3]) : "l"(q) : "memory");
          else if (LW == 2) asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(ln[r].w[0]), "=l"(ln[r].w[1]) : "l"(q) : "memory");
          else asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(ln[r].w[0]) : "l"(q) : "memory");
        }
        if (!found) {  // table refuses new keys: the row is replayed after the table has grown
          const unsigned long long at = atomicAdd(&p.counters[1], 1ull);
          p.ovf_rows[at] = (unsigned)(p.row_begin + i + r);
          continue;
        }
      }
      unsigned long long* const line = base + slot[r] * LW;
      if (HAS_MIN || HAS_MAX) {
        if (!is_nan_val(v[r], MT)) {  // f64::min / f64::max ignore NaN (aggregate.rs:139-140, 208-209)
          const unsigned long long e = ord_enc(v[r], MT);
          if (HAS_MIN && (!have_line || !p.cond_mm || e < line_word(ln[r], p.lean.min_w))) atomicMin(line + p.lean.min_w, e);
          if (HAS_MAX && (!have_line || !p.cond_mm || e > line_word(ln[r], p.lean.max_w))) atomicMax(line + p.lean.max_w, e);
        }
      }
      if (HAS_SUM) {
        if (MT == MT_F64) atomicAdd((double*)(p.lean.sum_arr + slot[r]), u2d(v[r]));
        else atomicAdd(p.lean.sum_arr + slot[r], v[r]);
      }
      if (HAS_CNT) atomicAdd(p.lean.cnt_arr + slot[r], 1ull);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) new_groups += __shfl_xor_sync(0xffffffffu, new_groups, o);
    if (lane == 0 && new_groups) atomicAdd(&p.counters[0], (unsigned long long)new_groups);
  }
}

// ---- wide / mixed keys ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long wide_tag(unsigned long long h) {  // ready form: low bit set, never EMPTY_KEY
  unsigned long long t = h | 1ull;
  if (t == EMPTY_KEY) t ^= 2ull;
  return t;
}
__device__ __forceinline__ bool utf8_equal(const int* off_a, const unsigned char* bytes_a, long long ra, const Utf8Source& sb, long long rb) {
  const int sa = off_a[ra], la = off_a[ra + 1] - sa, s2 = sb.off[rb], lb = sb.off[rb + 1] - s2;
  if (la != lb) return false;
  for (int i = 0; i < la; i++)
    if (bytes_a[sa + i] != sb.bytes[s2 + i]) return false;
  return true;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* q) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(q) : "memory");
  return v;
}

// K5 for wide keys.  Claim protocol of a slot: CAS the tag word EMPTY -> tag & ~1 (busy), store the key parts,
// fence, store tag | 1 (ready).  A row that meets a busy slot with its own tag polls briefly and otherwise
// defers itself to the replay list (counters[5] counts those: they need no table growth, the slot is
// ready by the time the replay runs), so no thread ever waits on another one indefinitely.
template <int DEPTH, bool NULLS>
__global__ void __launch_bounds__(AG_THREADS) k_hash_agg_wide(const __grid_constant__ AggParams p) {
  typedef InterpSrc<DEPTH, NULLS> Src;
  constexpr int R = Src::R;
  const int tid = threadIdx.x, lane = tid & 31;
  const long long n = p.row_list ? p.nlist : p.nrows;
  const int hshift = hash_shift(p.cap);
  const unsigned long long smask = (unsigned long long)p.cap - 1ull;
  const int KW = p.wide.kw;
  bool bad = false;
  constexpr int TILE = AG_THREADS * R;
  for (long long tb = (long long)blockIdx.x * TILE; tb < n; tb += (long long)gridDim.x * TILE) {
    unsigned long long filled = 0;
    if (lane == 0) filled = __ldcg(&p.counters[0]);
    filled = __shfl_sync(0xffffffffu, filled, 0);
    const bool full = (long long)filled >= p.max_groups;
    Src src;
    src.load(p, tb, n, tid, 0ull);
    src.prepare(p);
    // key parts: value of a fixed-width part; for a Utf8 part the program reads the string's 64-bit hash
    unsigned long long part[kMaxKeys][R];
    unsigned long long hsh[R];
#pragma unroll
    for (int r = 0; r < R; r++) hsh[r] = 0x9e3779b97f4a7c15ull;
    for (int k = 0; k < kMaxKeys; k++) {
      if (k >= KW) break;
      unsigned long long v[R];
      src.key(p, k, v);
#pragma unroll
      for (int r = 0; r < R; r++) {
        part[k][r] = v[r];
        hsh[r] = mix64(hsh[r] ^ (v[r] + 0x9e3779b97f4a7c15ull * (unsigned long long)(k + 1)));
      }
    }
    long long slot[R];
    unsigned new_groups = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      slot[r] = -1;
      if (!((src.mask >> r) & 1u)) continue;
      const long long row = src.g.rows[r];
      const unsigned long long ready = wide_tag(hsh[r]), busy = ready & ~1ull;
      unsigned long long h = home_slot(hsh[r], hshift);
      bool deferred = false;
      for (int probes = 0; probes < AG_MAX_PROBE; ++probes) {
        unsigned long long* line = p.t.key((long long)h);
        unsigned long long t = ld_volatile_u64(line);
        if (t == EMPTY_KEY) {
          if (full) break;
          t = atomicCAS(line, EMPTY_KEY, busy);
          if (t == EMPTY_KEY) {  // claimed: publish the key parts, then the ready tag
            for (int k = 0; k < KW; k++) line[1 + k] = p.wide.is_utf8[k] ? (p.wide.ref_base[k] | (unsigned long long)row) : part[k][r];
            __threadfence();
            atomicExch(line, ready);
            new_groups++;
            slot[r] = (long long)h;
            break;
          }
        }
        if ((t | 1ull) == ready) {
          for (int spin = 0; t == busy && spin < 64; spin++) { __nanosleep(64); t = ld_volatile_u64(line); }
          if (t == busy) { deferred = true; break; }
          __threadfence();
          bool same = true;
          for (int k = 0; same && k < KW; k++) {
            const unsigned long long w = __ldcg(line + 1 + k);
            if (p.wide.is_utf8[k]) {
              const Utf8Source& sc = p.wide.srcs[w >> UTF8_SRC_SHIFT];
              same = utf8_equal(p.wide.off[k], p.wide.bytes[k], row, sc, (long long)(w & ((1ull << UTF8_SRC_SHIFT) - 1ull)));
            } else {
              same = w == part[k][r];
            }
          }
          if (same) { slot[r] = (long long)h; break; }
        }
        h = (h + 1ull) & smask;
      }
      if (slot[r] < 0) {
        const unsigned long long at = atomicAdd(&p.counters[1], 1ull);
        p.ovf_rows[at] = (unsigned)row;
        if (deferred) atomicAdd(&p.counters[5], 1ull);
      }
    }
    for (int g = 0; g < p.nargs; g++) {
      unsigned long long v[R];
      unsigned av;
      const unsigned b = src.arg(p, g, v, av);
      for (int a = 0; a < p.naggs; a++) {
        if (p.agg_arg[a] != g) continue;
        const int func = p.aggs[a].func, mt = p.aggs[a].mtype;
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (NULLS && func == DFGPU_AGG_COUNT && !((av >> r) & 1u)) continue;
          if (slot[r] >= 0) {
            acc_fold_global(func, mt, p.t.val(slot[r], a), v[r]);
            if ((b >> r) & 1u) bad = true;
          }
        }
      }
    }
    bad = bad || src.bad != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) new_groups += __shfl_xor_sync(0xffffffffu, new_groups, o);
    if (lane == 0 && new_groups) atomicAdd(&p.counters[0], (unsigned long long)new_groups);
  }
  if (bad) p.counters[3] = 1ull;
}

// Re-insertion of the (distinct) groups of a wide-key table into a bigger one: every entry goes to the first
// empty slot after its home; no key comparison is needed because the source table holds each key once.
struct WideMoveParams {
  TableLayout from, to;
  long long from_cap, to_cap;
  int kw, naggs;
};
__global__ void __launch_bounds__(256) k_wide_move(const __grid_constant__ WideMoveParams p) {
  const int hshift = hash_shift(p.to_cap);
  const unsigned long long smask = (unsigned long long)p.to_cap - 1ull;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < p.from_cap; s += (long long)gridDim.x * blockDim.x) {
    const unsigned long long* src = p.from.key(s);
    const unsigned long long tag = src[0];
    if (tag == EMPTY_KEY) continue;
    unsigned long long h = home_slot(tag & ~1ull, hshift);  // the tag IS the hash (but for its low bit): same home rule as the scan
    for (;;) {
      unsigned long long* dst = p.to.key((long long)h);
      if (atomicCAS(dst, EMPTY_KEY, tag) == EMPTY_KEY) {
        for (int k = 0; k < p.kw; k++) dst[1 + k] = src[1 + k];
        for (int a = 0; a < p.naggs; a++) *p.to.val((long long)h, a) = *p.from.val(s, a);
        break;
      }
      h = (h + 1ull) & smask;
    }
  }
}

// K4: no GROUP BY.  Per-thread accumulators live in shared memory (one 8-byte cell per thread per
// aggregate, conflict-free), block-reduced at the end, one global reduction per CTA per aggregate.
constexpr int RD_R = 8;  // rows per thread per tile in the column reduce (more bytes in flight per SM)
constexpr int RD_TILE = AG_THREADS * RD_R;

// NULLS: array_ops::{min,max,sum} skip nulls and report None when nothing was non-null
// (restated from arrow 0.12; call sites aggregate.rs:347-541): the number of non-null inputs per
// aggregate is accumulated in counters[8 + a] so finish can emit a null.
template <int DEPTH, bool NULLS>
__global__ void __launch_bounds__(AG_THREADS) k_reduce(const __grid_constant__ AggParams p) {
  __shared__ unsigned long long s_acc[kMaxAggs][AG_THREADS];
  __shared__ unsigned long long s_red[AG_THREADS / 32];
  unsigned nn[kMaxAggs];
#pragma unroll
  for (int a = 0; a < kMaxAggs; a++) nn[a] = 0;
  unsigned npass = 0;  // rows that passed the fused predicate (counters[6]: an aggregate over zero rows is null)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int a = 0; a < p.naggs; a++) s_acc[a][tid] = agg_identity(p.aggs[a].func);
  bool bad = false;
  for (long long tb = (long long)blockIdx.x * RD_TILE; tb < p.nrows; tb += (long long)gridDim.x * RD_TILE) {
    GlobalRows<RD_R> src;
    src.valid = 0;
    long long (&rows)[RD_R] = src.rows;
#pragma unroll
    for (int r = 0; r < RD_R; r++) {
      const long long i = tb + (long long)r * AG_THREADS + tid;
      rows[r] = i < p.nrows ? i : -1;
      if (i < p.nrows) src.valid |= 1u << r;
    }
    unsigned rmask = src.valid;  // rows that pass the fused WHERE predicate (program 0)
    if (p.has_pred) {
      unsigned long long v[RD_R];
      unsigned pv;
      const unsigned b = eval_program_n<DEPTH, RD_R, false, NULLS>(p.ps, 0, src, v, pv);
      bad = bad || (b != 0);
      unsigned keep = 0;
#pragma unroll
      for (int r = 0; r < RD_R; r++) keep |= (unsigned)(v[r] & 1ull) << r;
      rmask &= keep;
    }
    if (p.has_pred) npass += __popc(rmask);
    for (int g = 0; g < p.nargs; g++) {
      unsigned long long v[RD_R];
      unsigned av;
      const unsigned b = eval_program_n<DEPTH, RD_R, false, NULLS>(p.ps, p.has_pred + g, src, v, av);
      bad = bad || ((b & rmask) != 0);
#pragma unroll
      for (int a = 0; a < kMaxAggs; a++) {
        if (a >= p.naggs || p.agg_arg[a] != g) continue;
        const int func = p.aggs[a].func, mt = p.aggs[a].mtype;
        unsigned long long acc = s_acc[a][tid];
#pragma unroll
        for (int r = 0; r < RD_R; r++)
          if (((rmask >> r) & 1u) && (!NULLS || ((av >> r) & 1u))) acc = acc_fold(func, mt, acc, v[r]);
        s_acc[a][tid] = acc;
        if (NULLS) nn[a] += __popc(av & rmask);
      }
    }
  }
  for (int a = 0; a < p.naggs; a++) {
    const int func = p.aggs[a].func, mt = p.aggs[a].mtype;
    unsigned long long acc = s_acc[a][tid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = acc_merge(func, mt, acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      acc = lane < AG_THREADS / 32 ? s_red[lane] : agg_identity(func);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc = acc_merge(func, mt, acc, __shfl_xor_sync(0xffffffffu, acc, o));
      if (lane == 0) acc_merge_global(func, mt, p.t.val(0, a), acc);  // cap == 0: slot 0
    }
    __syncthreads();
  }
  if (NULLS) {
#pragma unroll
    for (int a = 0; a < kMaxAggs; a++) {
      unsigned c = nn[a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if (lane == 0 && c && a < p.naggs) atomicAdd(&p.counters[8 + a], (unsigned long long)c);
    }
  }
  if (p.has_pred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) npass += __shfl_xor_sync(0xffffffffu, npass, o);
    if (lane == 0 && npass) atomicAdd(&p.counters[6], (unsigned long long)npass);
  }
  if (bad) p.counters[3] = 1ull;
}

// K4 fast path: every aggregate argument is a plain, null-free Float64 column.  One pass per distinct
// column with 128-bit loads; SUM / MIN / MAX / COUNT of the column are all cheap enough to be computed
// together and the requested ones are merged into the accumulators.
struct ReduceF64Params {
  const double* col[kMaxAggs];  // distinct argument columns
  int ncols;
  long long nrows;
  int naggs;
  AggDesc aggs[kMaxAggs];
  int agg_arg[kMaxAggs];
  TableLayout t;
};
__global__ void __launch_bounds__(256) k_reduce_f64(const __grid_constant__ ReduceF64Params p) {
  __shared__ unsigned long long s_red[3][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long npairs = p.nrows >> 1;
  for (int g = 0; g < p.ncols; g++) {
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(p.col[g]);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    unsigned long long mn = ~0ull, mx = 0ull;
    auto fold = [&](double v, double& s) {
      s += v;
      if (v == v) {  // MIN / MAX skip NaN (f64::min / f64::max, aggregate.rs:139-140,208-209)
        const unsigned long long e = ord_enc(d2u(v), MT_F64);
        mn = e < mn ? e : mn;
        mx = e > mx ? e : mx;
      }
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + tid;
    for (; i + stride < npairs; i += 2 * stride) {  // two independent 128-bit loads in flight
      const double2 a = __ldg(&c2[i]), b = __ldg(&c2[i + stride]);
      fold(a.x, s0); fold(a.y, s1); fold(b.x, s2); fold(b.y, s3);
    }
    if (i < npairs) { const double2 a = __ldg(&c2[i]); fold(a.x, s0); fold(a.y, s1); }
    if ((p.nrows & 1) && blockIdx.x == 0 && tid == 0) fold(p.col[g][p.nrows - 1], s2);
    unsigned long long sum = d2u((s0 + s1) + (s2 + s3));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum = d2u(u2d(sum) + u2d(__shfl_xor_sync(0xffffffffu, sum, o)));
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = a < mn ? a : mn;
      mx = b > mx ? b : mx;
    }
    if (lane == 0) { s_red[0][warp] = sum; s_red[1][warp] = mn; s_red[2][warp] = mx; }
    __syncthreads();
    if (warp == 0) {
      sum = lane < 8 ? s_red[0][lane] : 0ull;
      mn = lane < 8 ? s_red[1][lane] : ~0ull;
      mx = lane < 8 ? s_red[2][lane] : 0ull;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        sum = d2u(u2d(sum) + u2d(__shfl_xor_sync(0xffffffffu, sum, o)));
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
      }
      if (lane == 0) {
        for (int a = 0; a < p.naggs; a++) {
          if (p.agg_arg[a] != g) continue;
          const int f = p.aggs[a].func;
          if (f == DFGPU_AGG_SUM) atomicAdd((double*)p.t.val(0, a), u2d(sum));
          else if (f == DFGPU_AGG_MIN) atomicMin(p.t.val(0, a), mn);
          else if (f == DFGPU_AGG_MAX) atomicMax(p.t.val(0, a), mx);
          else if (blockIdx.x == 0) atomicAdd(p.t.val(0, a), (unsigned long long)p.nrows);  // COUNT of a null-free column
        }
      }
    }
    __syncthreads();
  }
}

// K7: scan the table, emit occupied slots densely.  raw != 0 keeps packed keys / undecoded
// accumulators (the exchange format of the multi-GPU merge).
struct CompactParams {
  TableLayout t;
  long long cap;
  int sentinel_used;
  int nkeys, naggs, raw;
  long long raw_stride;  // raw output: element idx of every raw array lives at idx * raw_stride (0 = 1: dense arrays)
  const unsigned long long* sentinel_flag;  // non-null: device flag that says whether the sentinel slot (slot cap) is in use
  int wide_kw;           // wide keys: line words 1..wide_kw are the key parts (a Utf8 part = a string reference)
  int key_is_utf8[kMaxKeys];
  AggDesc aggs[kMaxAggs];
  // aggregates over the same argument expression (MIN(v), MAX(v), SUM(v)) share one evaluation:
  // programs [nkeys, nkeys + nargs) are the DISTINCT argument programs, agg_arg[a] picks one
  int agg_arg[kMaxAggs];
  int nargs;
  unsigned long long key_mask[kMaxKeys];
  int key_shift[kMaxKeys];
  int key_dtype[kMaxKeys];
  void* out_keys[kMaxKeys];
  void* out_vals[kMaxAggs];
  unsigned long long* counter;
};

__global__ void __launch_bounds__(256) k_compact(const __grid_constant__ CompactParams p) {
  const int lane = threadIdx.x & 31;
  const long long total = p.cap + 1;
  const long long step = (long long)gridDim.x * blockDim.x;
  // round the loop bound up to a warp multiple so ballots stay converged
  for (long long s0 = (long long)blockIdx.x * blockDim.x; s0 < total; s0 += step) {
    const long long s = s0 + threadIdx.x;
    bool occ = false;
    unsigned long long key = 0;
    if (s < p.cap) { key = *p.t.key(s); occ = key != EMPTY_KEY; }
    else if (s == p.cap) { key = EMPTY_KEY; occ = p.sentinel_used != 0 || (p.sentinel_flag && *p.sentinel_flag != 0ull); }
    const unsigned m = __ballot_sync(0xffffffffu, occ);
    if (!m) continue;
    unsigned long long basei = 0;
    if (lane == 0) basei = atomicAdd(p.counter, (unsigned long long)__popc(m));
    basei = __shfl_sync(0xffffffffu, basei, 0);
    if (!occ) continue;
    const long long idx = (long long)(basei + __popc(m & ((1u << lane) - 1u)));
    if (p.raw) {
      const long long at = p.raw_stride ? idx * p.raw_stride : idx;
      ((unsigned long long*)p.out_keys[0])[at] = key;
      for (int a = 0; a < p.naggs; a++) ((unsigned long long*)p.out_vals[a])[at] = *p.t.val(s, a);
    } else {
      for (int k = 0; k < p.nkeys; k++) {
        if (p.wide_kw) {
          const unsigned long long w = p.t.key(s)[1 + k];
          if (p.key_is_utf8[k]) ((unsigned long long*)p.out_keys[k])[idx] = w;  // string reference, gathered by the host
          else store_elem(p.out_keys[k], p.key_dtype[k], idx, w);
          continue;
        }
        unsigned long long v = (key >> p.key_shift[k]) & p.key_mask[k];
        store_elem(p.out_keys[k], p.key_dtype[k], idx, v);
      }
      for (int a = 0; a < p.naggs; a++) {
        unsigned long long v = *p.t.val(s, a);
        const int f = p.aggs[a].func;
        if (f == DFGPU_AGG_MIN || f == DFGPU_AGG_MAX) v = ord_dec(v, p.aggs[a].mtype);
        store_elem(p.out_vals[a], p.aggs[a].out_dtype, idx, v);
      }
    }
  }
}

// Re-insert (raw key, raw accumulators) entries into a table: used to grow the table and to merge
// the partial aggregates of other GPUs (K6).
struct MergeParams {
  const unsigned long long* in_keys;
  const unsigned long long* in_vals[kMaxAggs];
  long long in_stride;  // entry i of every input array lives at i * in_stride (0 = 1: dense arrays)
  long long n;
  TableLayout t;
  long long cap;
  int naggs;
  AggDesc aggs[kMaxAggs];
  unsigned long long* counters;
};

__global__ void __launch_bounds__(256) k_merge(const __grid_constant__ MergeParams p) {
  const int hshift = hash_shift(p.cap);
  unsigned new_groups = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const long long at = p.in_stride ? i * p.in_stride : i;
    const unsigned long long key = p.in_keys[at];
    long long slot;
    if (key == EMPTY_KEY) {
      p.counters[2] = 1ull;
      slot = p.cap;
    } else {
      const unsigned long long h = home_slot(mix64(key), hshift);
      Line ln;
      load_line<false>(p.t, (long long)h, ln);
      slot = probe_insert<false>(p.t, p.cap, key, ln, h, false, new_groups);
      if (slot < 0) { p.counters[3] = 2ull; continue; }  // cannot happen: caller sizes the table
    }
    for (int a = 0; a < p.naggs; a++)
      acc_merge_global(p.aggs[a].func, p.aggs[a].mtype, p.t.val(slot, a), p.in_vals[a][at]);
  }
  if (new_groups) atomicAdd(&p.counters[0], (unsigned long long)new_groups);
}

// ---- owner-partitioned exchange of partial aggregates (multi-GPU merge) -----------------
// Every group key has one OWNER rank, a function of the key alone, so that each rank merges only 1/W of the
// keys: owner = bits of mix64(key) that the table slot does not use.
constexpr int AG_MAX_WORLD = 64;
__device__ __forceinline__ int owner_of(unsigned long long key, int world) {
  return key == EMPTY_KEY ? 0 : (int)((mix64(key) & 0xffffffffull) % (unsigned long long)world);
}
struct OwnerParams {
  const unsigned long long* keys;              // raw (packed) keys of the local groups
  const unsigned long long* vals[kMaxAggs];    // raw accumulators, dense arrays
  long long n;
  int world, naggs;
  unsigned long long* counts;                  // [world] entries per owner
  unsigned long long seg_off[AG_MAX_WORLD];    // scatter: first entry of owner o's segment in `rows`
  unsigned long long* cursor;                  // [world], zeroed
  unsigned long long* rows;                    // scatter output: entries of (1 + naggs) words, grouped by owner
};
__global__ void __launch_bounds__(256) k_owner_count(const __grid_constant__ OwnerParams p) {
  __shared__ unsigned s_cnt[AG_MAX_WORLD];
  if (threadIdx.x < AG_MAX_WORLD) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(&s_cnt[owner_of(p.keys[i], p.world)], 1u);
  __syncthreads();
  if (threadIdx.x < p.world && s_cnt[threadIdx.x]) atomicAdd(&p.counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_owner_scatter(const __grid_constant__ OwnerParams p) {
  const long long E = 1 + p.naggs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = p.keys[i];
    const int o = owner_of(key, p.world);
    const unsigned long long j = p.seg_off[o] + atomicAdd(&p.cursor[o], 1ull);
    unsigned long long* row = p.rows + j * E;
    row[0] = key;
    for (int a = 0; a < p.naggs; a++) row[1 + a] = p.vals[a][i];
  }
}

// raw rows (key, accumulators...) -> typed result columns: group columns first, then aggregates
// (aggregate.rs:890-949); the decode half of k_compact for rows that arrive from the exchange
struct DecodeParams {
  const unsigned long long* rows;
  long long n;
  // rows arrive as `nseg` segments of seg_stride entries each, of which the first seg_n[r] are real (the padded
  // all-gather of the owned groups); nseg == 0: one dense list
  int nseg;
  long long seg_stride;
  long long seg_n[AG_MAX_WORLD];
  int nkeys, naggs;
  AggDesc aggs[kMaxAggs];
  unsigned long long key_mask[kMaxKeys];
  int key_shift[kMaxKeys];
  int key_dtype[kMaxKeys];
  void* out_keys[kMaxKeys];
  void* out_vals[kMaxAggs];
};
__global__ void __launch_bounds__(256) k_decode_rows(const __grid_constant__ DecodeParams p) {
  const long long E = 1 + p.naggs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    long long at = i;
    if (p.nseg) {  // dense output index -> (segment, entry)
      long long j = i;
      int r = 0;
      while (r < p.nseg - 1 && j >= p.seg_n[r]) j -= p.seg_n[r++];
      at = (long long)r * p.seg_stride + j;
    }
    const unsigned long long* row = p.rows + at * E;
    const unsigned long long key = row[0];
    for (int k = 0; k < p.nkeys; k++) store_elem(p.out_keys[k], p.key_dtype[k], i, (key >> p.key_shift[k]) & p.key_mask[k]);
    for (int a = 0; a < p.naggs; a++) {
      unsigned long long v = row[1 + a];
      const int f = p.aggs[a].func;
      if (f == DFGPU_AGG_MIN || f == DFGPU_AGG_MAX) v = ord_dec(v, p.aggs[a].mtype);
      store_elem(p.out_vals[a], p.aggs[a].out_dtype, i, v);
    }
  }
}

// Utf8 GROUP BY keys are grouped by a 64-bit hash of the string; accumulator `rep_agg` holds the
// earliest (source << 40 | row) of each group.  Every row's string must equal its group's
// representative string, otherwise two different strings collided on the hash.
struct VerifyParams {
  const unsigned long long* hashes;
  long long n;
  TableLayout t;
  long long cap;
  int rep_agg;
  const int* off;
  const unsigned char* bytes;
  const Utf8Source* srcs;
  unsigned long long* flag;
};
__global__ void __launch_bounds__(256) k_utf8_group_verify(const __grid_constant__ VerifyParams p) {
  const unsigned long long hmask = (unsigned long long)p.cap - 1ull;
  const int hshift = hash_shift(p.cap);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = p.hashes[i];
    long long slot = p.cap;
    if (key != EMPTY_KEY) {
      unsigned long long h = home_slot(mix64(key), hshift);
      for (;;) {
        const unsigned long long cur = __ldcg(p.t.key((long long)h));
        if (cur == key) { slot = (long long)h; break; }
        if (cur == EMPTY_KEY) { slot = -1; break; }
        h = (h + 1ull) & hmask;
      }
    }
    if (slot < 0) { *p.flag = 2ull; continue; }
    const unsigned long long rep = *p.t.val(slot, p.rep_agg);
    const Utf8Source& sc = p.srcs[rep >> UTF8_SRC_SHIFT];
    const long long r = (long long)(rep & ((1ull << UTF8_SRC_SHIFT) - 1ull));
    const int s = p.off[i], len = p.off[i + 1] - s, s2 = sc.off[r], len2 = sc.off[r + 1] - s2;
    bool same = len == len2;
    for (int b = 0; same && b < len; b++) same = p.bytes[s + b] == sc.bytes[s2 + b];
    if (!same) *p.flag = 1ull;
  }
}

// fill the table with empty keys and accumulator identities
struct InitParams {
  TableLayout t;
  long long nslots;
  int naggs;
  unsigned long long ident[kMaxAggs];
};
__global__ void __launch_bounds__(256) k_table_init(const __grid_constant__ InitParams p) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < p.nslots; s += (long long)gridDim.x * blockDim.x) {
    *p.t.key(s) = EMPTY_KEY;
    for (int a = 0; a < p.naggs; a++) *p.t.val(s, a) = p.ident[a];
  }
}

}  // namespace dfgpu

using namespace dfgpu;

// ---------------------------------------------------------------------------------------------
// host state of one AggregateRelation
// ---------------------------------------------------------------------------------------------
struct dfgpu_aggstate {
  dfgpu_ctx* ctx = nullptr;
  std::vector<std::vector<dfgpu_insn>> key_progs;
  std::vector<std::vector<dfgpu_insn>> arg_progs;
  std::vector<int> funcs, out_dtypes;
  int nkeys = 0, naggs = 0;
  // resolved at the first update
  bool typed = false;
  std::vector<int> key_dtypes;
  std::vector<int> key_shift;
  std::vector<unsigned long long> key_mask;
  std::vector<AggDesc> descs;
  // table
  long long cap = 0;
  long long expected = 0;
  TableLayout t{};
  bool aos = false;        // "line" layout: every accumulator shares the line with the key (tables beyond L2)
  std::vector<dfgpu_insn> pred_prog;  // fused WHERE predicate (dfgpu_aggregate_set_predicate); empty = none
  bool use_front = false;  // route rows through the per-CTA shared-memory front table
  int npass = 1;           // scan passes per big batch (tables that outgrow L2, see AggParams)
  // wide keys: composite keys of more than 64 bits or with Utf8 parts (k_hash_agg_wide)
  bool wide = false;
  std::vector<int> key_is_utf8;
  // Utf8 GROUP BY key (aggregate.rs:842-847): grouped by a 64-bit string hash; the last accumulator is a
  // hidden MIN(source << 40 | row) = representative row of the group; the key columns of all batches are
  // retained so the representatives' strings can be gathered at finish
  bool utf8_key = false;
  std::vector<void*> utf8_owned;         // retained offsets / bytes buffers (device)
  std::vector<Utf8Source> utf8_srcs;     // host copy of the source table
  Utf8Source* d_utf8_srcs = nullptr;
  bool saw_nulls = false;  // some batch went through the null-aware reduce: per-aggregate non-null counts are on the device
  std::vector<long long> nonnull_host = std::vector<long long>(8, 0);
  unsigned long long* d_counters = nullptr;  // 8 x u64
  long long ngroups = 0;
  bool sentinel_used = false;
  long long rows_seen = 0;
  bool finished = false;

  ~dfgpu_aggstate() {
    if (ctx) {
      ctx->free(t.base);
      ctx->free(d_counters);
      for (void* p : utf8_owned) ctx->free(p);
      ctx->free(d_utf8_srcs);
    }
  }
};

namespace {

// DFGPU_TRACE=1: print host-side phase timings of the aggregate operator (each phase synchronised)
struct Trace {
  bool on;
  dfgpu_ctx* ctx;
  double t0;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  explicit Trace(dfgpu_ctx* c) : on(getenv("DFGPU_TRACE") != nullptr), ctx(c), t0(0) {
    if (on) { cudaStreamSynchronize(ctx->stream); t0 = now(); }
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    const double t = now();
    fprintf(stderr, "[dfgpu trace] %-28s %8.3f ms\n", what, t - t0);
    t0 = t;
  }
};

long long next_pow2(long long x) {
  long long p = 1;
  while (p < x) p <<= 1;
  return p;
}

int grid_for(dfgpu_ctx* ctx, long long work_items, int per_block, int blocks_per_sm);

// groups x (1 + naggs) sectors is what the SoA layout keeps hot in L2; beyond this many bytes the
// table is built AoS (one sector per group).  B200 L2 = 126 MB, shared with the streaming input.
constexpr long long AG_SOA_L2_BUDGET = 64ll << 20;

// Cardinality estimate from a prefix sample: `d` distinct keys among the first `s` rows.  Under a
// uniform model E[d] = G (1 - exp(-s / G)); solved for G by bisection and capped by the rows of the
// batch.  Skewed keys make this an under-estimate, which the growth path absorbs.
long long estimate_groups(long long d, long long s, long long total_rows) {
  if (d <= 0) return 0;
  if (double(d) >= 0.995 * double(s)) return total_rows;  // (almost) every sampled row was a new group
  double lo = double(d), hi = 1e13;
  for (int it = 0; it < 80; it++) {
    const double g = 0.5 * (lo + hi);
    const double e = g * -expm1(-double(s) / g);
    if (e < double(d)) lo = g; else hi = g;
  }
  const double g = 0.5 * (lo + hi);
  return g > double(total_rows) ? total_rows : (long long)g;
}

// Bytes the scan keeps hot in L2 with the hybrid layout: one sector (or lw/4) per group for the line, and
// every sector of each additive array once a quarter of its slots is in use.  Beyond the budget the table
// is built in "line" form (one sector per group, whatever the number of aggregates).
bool hybrid_enabled() {
  static const bool off = getenv("DFGPU_AGG_HYBRID") && atoi(getenv("DFGPU_AGG_HYBRID")) == 0;  // A/B switch: 0 = plain SoA (round-1 layout)
  return !off;
}
void layout_shape(const std::vector<AggDesc>& descs, int naggs, int nkeys, bool line_mode, long long* lw, int* n_add, signed char* loc, int kw = 0) {
  *lw = 1;
  *n_add = 0;
  if (nkeys > 0 && (line_mode || kw > 0)) {  // wide keys: tag, kw key parts, then every accumulator, all in the line
    while (*lw < 1 + kw + naggs) *lw <<= 1;
    for (int a = 0; a < naggs; a++) loc[a] = (signed char)(1 + kw + a);
    return;
  }
  int nmm = 0;
  for (int a = 0; a < naggs; a++) {
    const bool mm = descs[size_t(a)].func == DFGPU_AGG_MIN || descs[size_t(a)].func == DFGPU_AGG_MAX;
    if (nkeys > 0 && mm && hybrid_enabled()) loc[a] = (signed char)(1 + nmm++);
    else loc[a] = (signed char)~((*n_add)++);
  }
  while (*lw < 1 + nmm) *lw <<= 1;
}
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
bool want_aos(long long groups, const std::vector<AggDesc>& descs, int naggs) {
  static const char* force = getenv("DFGPU_AGG_LAYOUT");  // A/B switch: "line" | "hybrid"
  if (force && std::string(force) == "line") return true;
  if (force && std::string(force) == "hybrid") return false;
  long long lw;
  int n_add;
  signed char loc[kMaxAggs];
  layout_shape(descs, naggs, 1, false, &lw, &n_add, loc);
  const long long cap = std::max(AG_MIN_CAP, next_pow2(2 * groups));
  const long long line_hot = std::min(cap * 8 * lw, groups * std::max<long long>(32, 8 * lw));
  const long long arr_hot = std::min(cap * 8, groups * 32);
  static const long long budget = std::max<long long>(AG_SOA_L2_BUDGET, (long long)env_int("DFGPU_AGG_PASS_MB", 32) * env_int("DFGPU_AGG_MAX_PASSES", 1) << 20);
  return line_hot + n_add * arr_hot > budget;
}

TableLayout table_alloc(dfgpu_ctx* ctx, int naggs, int nkeys, const std::vector<AggDesc>& descs, long long cap, bool aos, int kw = 0) {
  InitParams ip;
  memset(&ip, 0, sizeof(ip));
  int n_add = 0;
  layout_shape(descs, naggs, nkeys, aos, &ip.t.lw, &n_add, ip.t.loc, kw);
  const size_t line_words = size_t(cap + 1) * size_t(ip.t.lw);
  const size_t words = line_words + size_t(n_add) * size_t(cap + 1);
  ip.t.base = (unsigned long long*)ctx->alloc(words * 8);  // >= 256-byte aligned: lines of up to 32 bytes never straddle a sector
  ip.t.add = ip.t.base + line_words;
  ip.t.astride = cap + 1;
  ip.nslots = cap + 1;
  ip.naggs = naggs;
  for (int a = 0; a < naggs; a++) ip.ident[a] = agg_identity(descs[size_t(a)].func);
  k_table_init<<<grid_for(ctx, ip.nslots, 256 * 4, 16), 256, 0, ctx->stream>>>(ip);
  DF_CUDA(cudaGetLastError());
  ctx->launches++;
  return ip.t;
}

void read_counters(dfgpu_aggstate* st, unsigned long long* host8) {
  dfgpu_ctx* ctx = st->ctx;
  DF_CUDA(cudaMemcpyAsync(ctx->h_scratch + 8, st->d_counters, 64, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 8; i++) host8[i] = ctx->h_scratch[8 + i];
}

int grid_for(dfgpu_ctx* ctx, long long work_items, int per_block, int blocks_per_sm) {
  long long g = (work_items + per_block - 1) / per_block;
  long long cap = (long long)ctx->sm_count * blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return int(g);
}

// grow the table to new_cap, re-inserting every occupied slot
void table_grow(dfgpu_aggstate* st, long long new_cap) {
  dfgpu_ctx* ctx = st->ctx;
  const bool aos = st->aos || want_aos(std::max(st->ngroups, new_cap / 8), st->descs, st->naggs);
  TableLayout nt = table_alloc(ctx, st->naggs, st->nkeys, st->descs, new_cap, aos);
  // compact raw, then merge into the new table
  const size_t cnt = size_t(st->ngroups + 1);
  unsigned long long* ck = (unsigned long long*)ctx->alloc(cnt * 8);
  unsigned long long* cv = (unsigned long long*)ctx->alloc(cnt * 8 * size_t(st->naggs > 0 ? st->naggs : 1));
  CompactParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.t = st->t;
  cp.cap = st->cap;
  cp.sentinel_used = st->sentinel_used;
  cp.nkeys = st->nkeys;
  cp.naggs = st->naggs;
  cp.raw = 1;
  cp.out_keys[0] = ck;
  for (int a = 0; a < st->naggs; a++) {
    cp.aggs[a] = st->descs[size_t(a)];
    cp.out_vals[a] = cv + size_t(a) * cnt;
  }
  DF_CUDA(cudaMemsetAsync(st->d_counters + 4, 0, 8, ctx->stream));
  cp.counter = st->d_counters + 4;
  k_compact<<<grid_for(ctx, st->cap + 1, 256, 8), 256, 0, ctx->stream>>>(cp);
  DF_CUDA(cudaGetLastError());
  ctx->launches++;
  MergeParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.in_keys = ck;
  for (int a = 0; a < st->naggs; a++) {
    mp.in_vals[a] = cv + size_t(a) * cnt;
    mp.aggs[a] = st->descs[size_t(a)];
  }
  mp.n = st->ngroups + (st->sentinel_used ? 1 : 0);
  mp.t = nt;
  mp.cap = new_cap;
  mp.naggs = st->naggs;
  DF_CUDA(cudaMemsetAsync(st->d_counters, 0, 8, ctx->stream));  // ngroups is recounted by the merge
  mp.counters = st->d_counters;
  if (mp.n > 0) {
    k_merge<<<grid_for(ctx, mp.n, 256, 8), 256, 0, ctx->stream>>>(mp);
    DF_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->free(ck);
  ctx->free(cv);
  ctx->free(st->t.base);
  st->t = nt;
  st->aos = aos;
  st->cap = new_cap;
}

// Hot bytes of the hybrid layout for `groups` groups (see want_aos) and the number of scan passes that
// keeps the part of the table one pass touches L2 resident.
long long hybrid_hot_bytes(long long groups, const std::vector<AggDesc>& descs, int naggs) {
  long long lw;
  int n_add;
  signed char loc[kMaxAggs];
  layout_shape(descs, naggs, 1, false, &lw, &n_add, loc);
  const long long cap = std::max(AG_MIN_CAP, next_pow2(2 * groups));
  return std::min(cap * 8 * lw, groups * std::max<long long>(32, 8 * lw)) + n_add * std::min(cap * 8, groups * 32);
}
int passes_for(long long groups, const std::vector<AggDesc>& descs, int naggs) {
  static const int forced = env_int("DFGPU_AGG_PASSES", 0);          // A/B switch: 1 | 2 | 4 | 8
  static const int pass_mb = env_int("DFGPU_AGG_PASS_MB", 32);       // hot megabytes one pass may touch
  static const int max_passes = env_int("DFGPU_AGG_MAX_PASSES", 1);  // measured (profiles/r02d_microbench_agg.txt): a second pass over the input costs what the L2 hits save; off by default
  if (forced > 0) return forced;
  const long long hot = hybrid_hot_bytes(groups, descs, naggs);
  int np = 1;
  while (np < max_passes && hot / np > ((long long)pass_mb << 20)) np <<= 1;
  return hot / np > ((long long)pass_mb << 20) ? 1 : np;  // beyond max_passes x pass_mb: one pass over the line layout
}

// Launch one scan kernel.  FRONT launches admit the keys of every CTA's front table unconditionally when
// the CTA retires, so the fill limit of the global path is lowered by what they can add (grid x front
// slots): the table stays at most half full and the front merge always finds a slot.
template <class Kern>
void launch_scan(dfgpu_ctx* ctx, Kern kern, AggParams& p, long long n, bool front) {
  const size_t smem = front ? size_t(AG_FRONT_SLOTS) * 8 * size_t(1 + p.naggs) : 0;
  if (front && ctx->first_use((const void*)kern))
    DF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AG_FRONT_SLOTS * 8 * (1 + kMaxAggs)));
  int per_sm = 0;
  DF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, AG_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  const int grid = grid_for(ctx, n, AG_TILE, per_sm);
  if (front) p.max_groups = std::max<long long>(0, p.max_groups - (long long)grid * AG_FRONT_SLOTS);
  const int ps = ctx->prof_begin();
  kern<<<grid, AG_THREADS, smem, ctx->stream>>>(p);
  DF_CUDA(cudaGetLastError());
  ctx->prof_end(ps);
  ctx->launches++;
}
template <int DEPTH>
void launch_hash_agg(dfgpu_ctx* ctx, AggParams& p, long long n, bool front) {
  if (front) launch_scan(ctx, k_hash_agg<DEPTH, true, false>, p, n, true);
  else launch_scan(ctx, k_hash_agg<DEPTH, false, false>, p, n, false);
}
template <int M>
void launch_lean_m(dfgpu_ctx* ctx, AggParams& p, long long n, int mt) {
  if (mt == MT_F64) launch_scan(ctx, k_hash_agg_lean<M, MT_F64>, p, n, false);
  else if (mt == MT_I) launch_scan(ctx, k_hash_agg_lean<M, MT_I>, p, n, false);
  else launch_scan(ctx, k_hash_agg_lean<M, MT_U>, p, n, false);
}
void launch_lean(dfgpu_ctx* ctx, AggParams& p, long long n, int mask, int mt) {
  switch (mask) {
#define DF_LEAN(M) case M: launch_lean_m<M>(ctx, p, n, mt); break;
    DF_LEAN(1) DF_LEAN(2) DF_LEAN(3) DF_LEAN(4) DF_LEAN(5) DF_LEAN(6) DF_LEAN(7) DF_LEAN(8)
    DF_LEAN(9) DF_LEAN(10) DF_LEAN(11) DF_LEAN(12) DF_LEAN(13) DF_LEAN(14) DF_LEAN(15)
#undef DF_LEAN
    default: fail(DFGPU_ERR_INTERNAL, "lean kernel: bad aggregate mask");
  }
}
template <int DEPTH, bool NULLS = false>
void launch_reduce(dfgpu_ctx* ctx, const AggParams& p, long long n) {
  int per_sm = 0;
  DF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_reduce<DEPTH, NULLS>, AG_THREADS, 0));
  if (per_sm < 1) per_sm = 1;
  const int ps = ctx->prof_begin();
  k_reduce<DEPTH, NULLS><<<grid_for(ctx, n, RD_TILE, per_sm), AG_THREADS, 0, ctx->stream>>>(p);
  DF_CUDA(cudaGetLastError());
  ctx->prof_end(ps);
  ctx->launches++;
}

}  // namespace

extern "C" int dfgpu_aggregate_create(dfgpu_ctx* ctx, const dfgpu_insn* const* keys, const int* key_len, int nkeys,
                                      const dfgpu_agg* aggs, int naggs, int64_t expected_groups, dfgpu_aggstate** out) {
  return guarded([&] {
    if (!ctx || !out) fail(DFGPU_ERR_GENERAL, "dfgpu_aggregate_create: null argument");
    if (nkeys < 0 || nkeys > kMaxKeys) fail(DFGPU_ERR_NOT_IMPLEMENTED, "more than " + std::to_string(kMaxKeys) + " GROUP BY expressions");
    if (naggs < 1) fail(DFGPU_ERR_GENERAL, "aggregate needs at least one aggregate expression");
    if (naggs > kMaxAggs) fail(DFGPU_ERR_NOT_IMPLEMENTED, "more than " + std::to_string(kMaxAggs) + " aggregate expressions");
    ctx->use();
    auto st = std::make_unique<dfgpu_aggstate>();
    st->ctx = ctx;
    st->nkeys = nkeys;
    st->naggs = naggs;
    st->expected = expected_groups;
    for (int k = 0; k < nkeys; k++) st->key_progs.emplace_back(keys[k], keys[k] + key_len[k]);
    for (int a = 0; a < naggs; a++) {
      // compile_expr accepts min/max/count/sum (expression.rs:98-107); anything else is
      // General("Unsupported aggregate function ...")
      if (aggs[a].func < DFGPU_AGG_MIN || aggs[a].func > DFGPU_AGG_COUNT)
        fail(DFGPU_ERR_GENERAL, "Unsupported aggregate function '" + std::to_string(aggs[a].func) + "'");
      st->arg_progs.emplace_back(aggs[a].arg, aggs[a].arg + aggs[a].arg_len);
      st->funcs.push_back(aggs[a].func);
      st->out_dtypes.push_back(aggs[a].out_dtype);
    }
    st->d_counters = (unsigned long long*)ctx->alloc(128);  // [0..7] see AggParams, [8..15] non-null inputs per aggregate
    DF_CUDA(cudaMemsetAsync(st->d_counters, 0, 128, ctx->stream));
    *out = st.release();
  });
}

extern "C" int dfgpu_aggregate_set_predicate(dfgpu_aggstate* st, const dfgpu_insn* pred, int pred_len) {
  return guarded([&] {
    if (!st || (pred_len > 0 && !pred)) fail(DFGPU_ERR_GENERAL, "dfgpu_aggregate_set_predicate: null argument");
    if (st->rows_seen > 0 || st->typed) fail(DFGPU_ERR_GENERAL, "the predicate must be set before the first batch");
    st->pred_prog.assign(pred, pred + (pred_len > 0 ? pred_len : 0));
  });
}

namespace {
void agg_update(dfgpu_aggstate* st, const dfgpu_batch* batch);
}

extern "C" int dfgpu_aggregate_update(dfgpu_aggstate* st, const dfgpu_batch* batch) {
  return guarded([&] {
    if (!st || !batch) fail(DFGPU_ERR_GENERAL, "dfgpu_aggregate_update: null argument");
    agg_update(st, batch);
  });
}

// One big host RecordBatch: row-range chunks, every H2D copy queued up front on the copy-in stream, the
// scan of chunk c waits only for chunk c's copies — PCIe and the scan kernel overlap, and the table is
// the only state carried from chunk to chunk (update_accumulators is per row: aggregate.rs:548-612).
extern "C" int dfgpu_aggregate_update_host(dfgpu_aggstate* st, const dfgpu_col* cols, int ncols, int64_t chunk_rows) {
  return guarded([&] {
    if (!st || (ncols > 0 && !cols)) fail(DFGPU_ERR_GENERAL, "dfgpu_aggregate_update_host: null argument");
    dfgpu_ctx* ctx = st->ctx;
    ctx->use();
    const int64_t n = ncols > 0 ? cols[0].len : 0;
    bool streamable = n > 0;
    for (int i = 0; i < ncols; i++) {
      if (cols[i].len != n) fail(DFGPU_ERR_GENERAL, "all columns of a RecordBatch must have the same length");
      streamable = streamable && dtype_width(cols[i].dtype) > 0 && !cols[i].validity && cols[i].values;
    }
    if (chunk_rows <= 0) chunk_rows = 8ll << 20;
    chunk_rows = (chunk_rows + 1023) / 1024 * 1024;  // chunks start on even, 16-byte aligned rows of every column
    if (!streamable || n <= chunk_rows) {
      // small, nullable or variable-width batches: plain upload (Utf8 / validity handling lives there)
      dfgpu_batch* b = nullptr;
      int rc = dfgpu_batch_upload(ctx, cols, ncols, &b);
      if (rc != DFGPU_OK) fail(rc, dfgpu_last_error());
      struct G { dfgpu_batch* b; ~G() { dfgpu_batch_free(b); } } g{b};
      agg_update(st, b);
      return;
    }
    const int nchunks = int((n + chunk_rows - 1) / chunk_rows);
    std::vector<void*> dev(size_t(ncols), nullptr);
    std::vector<cudaEvent_t> evs(size_t(nchunks), nullptr);
    struct Cleanup {
      dfgpu_ctx* c; std::vector<void*>* d; std::vector<cudaEvent_t>* e;
      ~Cleanup() {
        cudaStreamSynchronize(c->stream_in);
        cudaStreamSynchronize(c->stream);
        for (void* q : *d) c->free(q);
        for (cudaEvent_t x : *e) if (x) cudaEventDestroy(x);
      }
    } cleanup{ctx, &dev, &evs};
    for (int i = 0; i < ncols; i++) dev[size_t(i)] = ctx->alloc(size_t(n) * size_t(dtype_width(cols[i].dtype)));
    // the device buffers may have been used on ctx->stream before (cached blocks): order the copies after it
    cudaEvent_t ready;
    DF_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    DF_CUDA(cudaEventRecord(ready, ctx->stream));
    DF_CUDA(cudaStreamWaitEvent(ctx->stream_in, ready, 0));
    DF_CUDA(cudaEventDestroy(ready));
    for (int c = 0; c < nchunks; c++) {
      const int64_t lo = int64_t(c) * chunk_rows, cnt = std::min<int64_t>(chunk_rows, n - lo);
      for (int i = 0; i < ncols; i++) {
        const size_t w = size_t(dtype_width(cols[i].dtype));
        DF_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(dev[size_t(i)]) + size_t(lo) * w,
                                static_cast<const uint8_t*>(cols[i].values) + size_t(cols[i].offset + lo) * w, size_t(cnt) * w,
                                cudaMemcpyHostToDevice, ctx->stream_in));
      }
      DF_CUDA(cudaEventCreateWithFlags(&evs[size_t(c)], cudaEventDisableTiming));
      DF_CUDA(cudaEventRecord(evs[size_t(c)], ctx->stream_in));
    }
    for (int c = 0; c < nchunks; c++) {
      const int64_t lo = int64_t(c) * chunk_rows, cnt = std::min<int64_t>(chunk_rows, n - lo);
      DF_CUDA(cudaStreamWaitEvent(ctx->stream, evs[size_t(c)], 0));
      dfgpu_batch view;  // borrows the chunk's slices of the device buffers
      view.ctx = ctx;
      view.owns = false;
      view.nrows = cnt;
      for (int i = 0; i < ncols; i++) {
        DevColumn d;
        d.dtype = cols[i].dtype;
        const size_t w = size_t(dtype_width(cols[i].dtype));
        d.values = static_cast<uint8_t*>(dev[size_t(i)]) + size_t(lo) * w;
        d.values_bytes = size_t(cnt) * w;
        view.cols.push_back(d);
      }
      agg_update(st, &view);
    }
  });
}

namespace {
void agg_update(dfgpu_aggstate* st, const dfgpu_batch* batch) {
  {
    if (st->finished) fail(DFGPU_ERR_GENERAL, "aggregate already finished");
    dfgpu_ctx* ctx = st->ctx;
    ctx->use();
    if (batch->nrows >= (1ll << 32)) fail(DFGPU_ERR_NOT_IMPLEMENTED, "batches of 2^32 rows or more");

    Trace tr(ctx);
    AggParams p;
    memset(&p, 0, sizeof(p));
    ProgramBuilder pb(batch);
    std::vector<int> kdt;
    // fused WHERE: program 0 (FilterRelation under the aggregate, context.rs:126-139)
    const int has_pred = st->pred_prog.empty() ? 0 : 1;
    if (has_pred) {
      const int pi = pb.add(st->pred_prog.data(), int(st->pred_prog.size()), "filter expression");
      if (pb.out_dtype(pi) != DFGPU_BOOL) fail(DFGPU_ERR_EXECUTION, "Filter expression did not evaluate to boolean");  // filter.rs:62-67
    }
    // a single plain Utf8 column as the key: group by a 64-bit hash of the strings (see dfgpu_aggstate)
    const DevColumn* ukey = nullptr;
    if (st->nkeys == 1 && st->key_progs[0].size() == 1 && st->key_progs[0][0].op == DFGPU_OP_COL && st->key_progs[0][0].col >= 0 &&
        size_t(st->key_progs[0][0].col) < batch->cols.size() && batch->cols[size_t(st->key_progs[0][0].col)].dtype == DFGPU_UTF8)
      ukey = &batch->cols[size_t(st->key_progs[0][0].col)];
    unsigned long long* d_hash = nullptr;
    struct HashFree { dfgpu_ctx* c; unsigned long long** p; ~HashFree() { c->free(*p); } } hash_free{ctx, &d_hash};
    std::vector<const DevColumn*> wide_ucols;    // per key part: its Utf8 column in this batch (or null)
    std::vector<unsigned long long*> wide_hashes;  // string hashes of the Utf8 parts
    struct HashesFree { dfgpu_ctx* c; std::vector<unsigned long long*>* v; ~HashesFree() { for (auto* q : *v) c->free(q); } } hashes_free{ctx, &wide_hashes};
    if (ukey) {
      if (st->typed && !st->utf8_key) fail(DFGPU_ERR_GENERAL, "GROUP BY key types changed between batches");
      if (!st->typed && st->naggs >= kMaxAggs) fail(DFGPU_ERR_NOT_IMPLEMENTED, "Utf8 GROUP BY key with " + std::to_string(kMaxAggs) + " aggregates");
      if (st->utf8_srcs.size() >= (1u << 20)) fail(DFGPU_ERR_NOT_IMPLEMENTED, "more than 2^20 batches with a Utf8 GROUP BY key");
      d_hash = (unsigned long long*)ctx->alloc(size_t(batch->nrows > 0 ? batch->nrows : 1) * 8);
      utf8_hash(ctx, *ukey, batch->nrows, d_hash);
      pb.add_synthetic_column(d_hash, DFGPU_UINT64);
      kdt.push_back(DFGPU_UINT64);
    } else {
      // key parts: integer expressions, and plain Utf8 columns (hashed here; the scan reads the hash as a column)
      for (int k = 0; k < st->nkeys; k++) {
        const auto& kp = st->key_progs[size_t(k)];
        const DevColumn* uc = nullptr;
        if (kp.size() == 1 && kp[0].op == DFGPU_OP_COL && kp[0].col >= 0 && size_t(kp[0].col) < batch->cols.size() &&
            batch->cols[size_t(kp[0].col)].dtype == DFGPU_UTF8)
          uc = &batch->cols[size_t(kp[0].col)];
        wide_ucols.push_back(uc);
        if (uc) {
          unsigned long long* dh = (unsigned long long*)ctx->alloc(size_t(batch->nrows > 0 ? batch->nrows : 1) * 8);
          wide_hashes.push_back(dh);
          utf8_hash(ctx, *uc, batch->nrows, dh);
          pb.add_synthetic_column(dh, DFGPU_UINT64);
          kdt.push_back(DFGPU_UTF8);
          continue;
        }
        int pi = pb.add(kp.data(), int(kp.size()), "GROUP BY expression");
        int dt = pb.out_dtype(pi);
        if (dt == DFGPU_UTF8) fail(DFGPU_ERR_NOT_IMPLEMENTED, "Utf8 GROUP BY keys must be plain columns");
        if (!is_int(dt)) fail(DFGPU_ERR_EXECUTION, "Unsupported GROUP BY data type");  // aggregate.rs:848-850
        kdt.push_back(dt);
      }
    }
    const int user_aggs = st->utf8_key ? st->naggs - 1 : st->naggs;  // the hidden representative is appended below
    std::vector<AggDesc> descs;
    std::vector<int> agg_arg(size_t(user_aggs), 0);
    int nargs = 0;
    for (int a = 0; a < user_aggs; a++) {
      // identical argument expressions are compiled (and evaluated) once
      int same = -1;
      for (int b = 0; b < a && same < 0; b++) {
        const auto &x = st->arg_progs[size_t(a)], &y = st->arg_progs[size_t(b)];
        if (x.size() == y.size() && memcmp(x.data(), y.data(), x.size() * sizeof(dfgpu_insn)) == 0) same = b;
      }
      int pi;
      if (same >= 0) {
        agg_arg[size_t(a)] = agg_arg[size_t(same)];
        pi = has_pred + st->nkeys + agg_arg[size_t(a)];
      } else {
        pi = pb.add(st->arg_progs[size_t(a)].data(), int(st->arg_progs[size_t(a)].size()), "aggregate argument");
        agg_arg[size_t(a)] = nargs++;
      }
      int dt = pb.out_dtype(pi);
      if (!is_numeric(dt)) fail(DFGPU_ERR_EXECUTION, std::string("Unsupported data type for aggregate: ") + dtype_name(dt));
      AggDesc d;
      d.func = uint8_t(st->funcs[size_t(a)]);
      d.mtype = mtype_of(dt);
      d.dtype = uint8_t(dt);
      int want = d.func == DFGPU_AGG_COUNT ? DFGPU_UINT64 : dt;
      int odt = st->out_dtypes[size_t(a)];
      if (odt == 0) odt = want;
      if (odt != want)  // the reference would hit "unexpected type when creating array from aggregate map" (aggregate.rs:683-695)
        fail(DFGPU_ERR_EXECUTION, "unexpected type when creating array from aggregate map");
      d.out_dtype = uint8_t(odt);
      descs.push_back(d);
    }
    if (ukey) {
      // hidden accumulator: MIN(source << 40 | row)
      pb.add_rowid_plus((unsigned long long)st->utf8_srcs.size() << UTF8_SRC_SHIFT);
      agg_arg.push_back(nargs++);
      AggDesc d;
      d.func = DFGPU_AGG_MIN;
      d.mtype = MT_U;
      d.dtype = DFGPU_UINT64;
      d.out_dtype = DFGPU_UINT64;
      descs.push_back(d);
      if (!st->typed) {
        st->utf8_key = true;
        st->naggs += 1;
      }
      // retain this batch's key column: representatives are gathered from it at finish
      const size_t ob = size_t(batch->nrows + 1) * 4, vb = ukey->values_bytes ? ukey->values_bytes : 1;
      int* off = (int*)ctx->alloc(ob);
      unsigned char* bytes = (unsigned char*)ctx->alloc(vb);
      DF_CUDA(cudaMemcpyAsync(off, ukey->offsets, ob, cudaMemcpyDeviceToDevice, ctx->stream));
      if (ukey->values_bytes) DF_CUDA(cudaMemcpyAsync(bytes, ukey->values, ukey->values_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      st->utf8_owned.push_back(off);
      st->utf8_owned.push_back(bytes);
      st->utf8_srcs.push_back(Utf8Source{off, bytes});
      ctx->free(st->d_utf8_srcs);
      st->d_utf8_srcs = (Utf8Source*)ctx->alloc(st->utf8_srcs.size() * sizeof(Utf8Source));
      DF_CUDA(cudaMemcpyAsync(st->d_utf8_srcs, st->utf8_srcs.data(), st->utf8_srcs.size() * sizeof(Utf8Source), cudaMemcpyHostToDevice, ctx->stream));
      DF_CUDA(cudaStreamSynchronize(ctx->stream));  // utf8_srcs may reallocate on the next batch
    }
    if (!st->typed) {
      st->key_dtypes = kdt;
      st->descs = descs;
      int bits = 0;
      bool any_utf8 = false;
      for (int k = st->nkeys - 1; k >= 0; k--) {  // last key in the low bits
        const bool u = !ukey && kdt[size_t(k)] == DFGPU_UTF8;
        any_utf8 = any_utf8 || u;
        int w = u ? 64 : dtype_width(kdt[size_t(k)]) * 8;
        st->key_shift.insert(st->key_shift.begin(), bits > 63 ? 0 : bits);
        st->key_mask.insert(st->key_mask.begin(), w == 64 ? ~0ull : ((1ull << w) - 1ull));
        st->key_is_utf8.insert(st->key_is_utf8.begin(), u ? 1 : 0);
        bits += w;
      }
      // more than 64 key bits, or Utf8 parts next to other parts: tagged slots with full key comparison
      st->wide = !ukey && (bits > 64 || any_utf8);
      if (st->nkeys == 1) st->key_mask[0] = ~0ull;  // single key: keep the sign-extended 64-bit value
      st->typed = true;
      st->cap = st->nkeys == 0 ? 0 : std::max(AG_MIN_CAP, next_pow2(2 * st->expected));
      if (st->wide) {
        st->aos = true;
        st->t = table_alloc(ctx, st->naggs, st->nkeys, st->descs, st->cap, true, st->nkeys);
      } else {
      st->aos = st->nkeys > 0 && want_aos(st->expected, st->descs, st->naggs);
      st->npass = (st->nkeys > 0 && !st->aos && st->expected > 0) ? passes_for(st->expected, st->descs, st->naggs) : 1;
      st->t = table_alloc(ctx, st->naggs, st->nkeys, st->descs, st->cap, st->aos);
      }
    } else {
      if (kdt != st->key_dtypes) fail(DFGPU_ERR_GENERAL, "GROUP BY key types changed between batches");
      for (int a = 0; a < st->naggs; a++)
        if (descs[size_t(a)].dtype != st->descs[size_t(a)].dtype) fail(DFGPU_ERR_GENERAL, "aggregate argument types changed between batches");
    }
    tr.mark("programs + table alloc");
    pb.finish(&p.ps);
    for (int s = 0; s < p.ps.ncols; s++)
      if (!is_numeric(p.ps.cols[s].dtype) && p.ps.cols[s].dtype != DFGPU_BOOL)  // Boolean columns: WHERE operands (boolean_ops!)
        fail(DFGPU_ERR_NOT_IMPLEMENTED, std::string("expressions over ") + dtype_name(p.ps.cols[s].dtype) + " columns are not supported on the GPU path yet");
    if (p.ps.max_depth > 8) fail(DFGPU_ERR_NOT_IMPLEMENTED, "expression too deep (register stack depth > 8)");
    st->rows_seen += batch->nrows;
    if (batch->nrows == 0) return;

    p.nkeys = st->nkeys;
    p.naggs = st->naggs;
    for (int a = 0; a < st->naggs; a++) {
      p.aggs[a] = st->descs[size_t(a)];
      p.agg_arg[a] = agg_arg[size_t(a)];
    }
    p.nargs = nargs;
    for (int k = 0; k < st->nkeys; k++) {
      p.key_mask[k] = st->key_mask[size_t(k)];
      p.key_shift[k] = st->key_shift[size_t(k)];
    }
    p.nrows = batch->nrows;
    p.counters = st->d_counters;
    p.has_pred = has_pred;
    {
      static const bool no_cond = getenv("DFGPU_AGG_COND_MM") && atoi(getenv("DFGPU_AGG_COND_MM")) == 0;  // A/B switch
      p.cond_mm = no_cond ? 0 : 1;
      static const bool hint = getenv("DFGPU_AGG_TABLE_HINT") && atoi(getenv("DFGPU_AGG_TABLE_HINT")) != 0;  // A/B switch (default off until measured)
      p.table_hint = hint ? 1 : 0;
    }
    const int d = p.ps.max_depth;

    if (st->nkeys == 0) {
      p.t = st->t;
      p.cap = 0;
      if (has_pred) DF_CUDA(cudaMemsetAsync(st->d_counters + 6, 0, 8, ctx->stream));  // rows passing the predicate
      if (p.ps.has_nulls) {
        st->saw_nulls = true;
        launch_reduce<8, true>(ctx, p, p.nrows);
      } else {
        if (!has_pred)
          for (int a = 0; a < st->naggs; a++) st->nonnull_host[size_t(a)] += batch->nrows;
        // fast path: every distinct argument is a plain Float64 column
        bool plain = !has_pred;
        ReduceF64Params rp;
        memset(&rp, 0, sizeof(rp));
        for (int g = 0; g < nargs && plain; g++) {
          const CompiledProgram& cpg = pb.prog(g);
          plain = cpg.is_plain_column && p.ps.cols[cpg.plain_slot].dtype == DFGPU_FLOAT64 &&
                  (reinterpret_cast<uintptr_t>(p.ps.cols[cpg.plain_slot].ptr) & 15) == 0;
          if (plain) rp.col[g] = (const double*)p.ps.cols[cpg.plain_slot].ptr;
        }
        if (plain) {
          rp.ncols = nargs;
          rp.nrows = p.nrows;
          rp.naggs = st->naggs;
          for (int a = 0; a < st->naggs; a++) { rp.aggs[a] = p.aggs[a]; rp.agg_arg[a] = p.agg_arg[a]; }
          rp.t = st->t;
          const int ps = ctx->prof_begin();
          k_reduce_f64<<<grid_for(ctx, p.nrows, 256 * 8, 8), 256, 0, ctx->stream>>>(rp);
          DF_CUDA(cudaGetLastError());
          ctx->prof_end(ps);
          ctx->launches++;
        } else if (d <= 1) launch_reduce<1>(ctx, p, p.nrows);
        else if (d <= 2) launch_reduce<2>(ctx, p, p.nrows);
        else if (d <= 4) launch_reduce<4>(ctx, p, p.nrows);
        else launch_reduce<8>(ctx, p, p.nrows);
      }
      unsigned long long c[8];
      read_counters(st, c);
      if (c[3]) fail(DFGPU_ERR_ARROW, "DivideByZero");
      if (has_pred && !p.ps.has_nulls)  // null-free inputs: every row that passed the predicate is a non-null input
        for (int a = 0; a < st->naggs; a++) st->nonnull_host[size_t(a)] += (long long)c[6];
      return;
    }

    if (st->wide) {
      // retain this batch's Utf8 key columns: a group's string lives in the batch whose row created it
      p.wide.kw = st->nkeys;
      bool new_src = false;
      for (int k = 0; k < st->nkeys; k++) {
        p.wide.is_utf8[k] = st->key_is_utf8[size_t(k)];
        const DevColumn* uc = wide_ucols[size_t(k)];
        if (!uc) continue;
        if (st->utf8_srcs.size() >= (1u << 20)) fail(DFGPU_ERR_NOT_IMPLEMENTED, "more than 2^20 retained Utf8 GROUP BY key columns");
        const size_t ob = size_t(batch->nrows + 1) * 4, vb = uc->values_bytes ? uc->values_bytes : 1;
        int* off = (int*)ctx->alloc(ob);
        unsigned char* bytes = (unsigned char*)ctx->alloc(vb);
        DF_CUDA(cudaMemcpyAsync(off, uc->offsets, ob, cudaMemcpyDeviceToDevice, ctx->stream));
        if (uc->values_bytes) DF_CUDA(cudaMemcpyAsync(bytes, uc->values, uc->values_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        p.wide.off[k] = off;
        p.wide.bytes[k] = bytes;
        p.wide.ref_base[k] = (unsigned long long)st->utf8_srcs.size() << UTF8_SRC_SHIFT;
        st->utf8_owned.push_back(off);
        st->utf8_owned.push_back(bytes);
        st->utf8_srcs.push_back(Utf8Source{off, bytes});
        new_src = true;
      }
      if (new_src) {
        ctx->free(st->d_utf8_srcs);
        st->d_utf8_srcs = (Utf8Source*)ctx->alloc(st->utf8_srcs.size() * sizeof(Utf8Source));
        DF_CUDA(cudaMemcpyAsync(st->d_utf8_srcs, st->utf8_srcs.data(), st->utf8_srcs.size() * sizeof(Utf8Source), cudaMemcpyHostToDevice, ctx->stream));
        DF_CUDA(cudaStreamSynchronize(ctx->stream));  // utf8_srcs may reallocate on the next batch
      }
      p.wide.srcs = st->d_utf8_srcs;
      unsigned* wovf[2] = {(unsigned*)ctx->alloc(size_t(batch->nrows) * 4), nullptr};
      struct WFreer { dfgpu_ctx* c; unsigned** o; ~WFreer() { c->free(o[0]); c->free(o[1]); } } wfreer{ctx, wovf};
      auto wide_grow = [&](long long new_cap) {
        TableLayout nt = table_alloc(ctx, st->naggs, st->nkeys, st->descs, new_cap, true, st->nkeys);
        WideMoveParams mp;
        memset(&mp, 0, sizeof(mp));
        mp.from = st->t;
        mp.to = nt;
        mp.from_cap = st->cap;
        mp.to_cap = new_cap;
        mp.kw = st->nkeys;
        mp.naggs = st->naggs;
        k_wide_move<<<grid_for(ctx, st->cap, 256, 8), 256, 0, ctx->stream>>>(mp);
        DF_CUDA(cudaGetLastError());
        ctx->launches++;
        DF_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->free(st->t.base);
        st->t = nt;
        st->cap = new_cap;
      };
      int cur = 0;
      const unsigned* list = nullptr;
      long long nlist = 0;
      for (int round = 0;; round++) {
        if (round > 60) fail(DFGPU_ERR_INTERNAL, "hash table growth did not converge");
        p.t = st->t;
        p.cap = st->cap;
        p.max_groups = st->cap / 2;
        p.row_begin = 0;
        p.nrows = batch->nrows;
        p.row_list = list;
        p.nlist = nlist;
        p.ovf_rows = wovf[cur];
        p.npass = 1;
        DF_CUDA(cudaMemsetAsync(st->d_counters + 1, 0, 8, ctx->stream));
        DF_CUDA(cudaMemsetAsync(st->d_counters + 5, 0, 8, ctx->stream));
        const long long n = list ? nlist : p.nrows;
        if (p.ps.has_nulls) launch_scan(ctx, k_hash_agg_wide<8, true>, p, n, false);
        else launch_scan(ctx, k_hash_agg_wide<8, false>, p, n, false);
        unsigned long long c[8];
        read_counters(st, c);
        if (c[3]) fail(DFGPU_ERR_ARROW, "DivideByZero");
        st->ngroups = (long long)c[0];
        const long long novf = (long long)c[1], deferred = (long long)c[5];
        tr.mark("scan kernel (wide keys)");
        if (novf == 0) {
          if (st->ngroups > st->cap / 2) wide_grow(st->cap * 4);
          break;
        }
        // rows that only met a slot still being published need no bigger table: replay them as they are
        if (!(novf == deferred && st->ngroups <= st->cap / 2)) wide_grow(st->cap * 4);
        list = wovf[cur];
        nlist = novf;
        cur ^= 1;
        if (!wovf[cur]) wovf[cur] = (unsigned*)ctx->alloc(size_t(batch->nrows) * 4);
      }
      return;
    }

    // Plain-column fast path: keys and arguments are plain 4/8-byte columns, the WHERE clause (if any) a
    // chain of column comparisons -> the interpreter-free kernel with 128-bit loads.
    bool use_plain = false;
    {
      static const bool off = getenv("DFGPU_AGG_PLAIN") && atoi(getenv("DFGPU_AGG_PLAIN")) == 0;  // A/B switch
      PlainSpec sp;
      memset(&sp, 0, sizeof(sp));
      auto wide = [](int dt) {
        return dt == DFGPU_FLOAT64 || dt == DFGPU_INT64 || dt == DFGPU_UINT64 || dt == DFGPU_FLOAT32 || dt == DFGPU_INT32 || dt == DFGPU_UINT32;
      };
      bool ok = !off && !p.ps.has_nulls && p.ps.ncols <= 4;
      for (int c = 0; ok && c < p.ps.ncols; c++)
        ok = wide(p.ps.cols[c].dtype) && (reinterpret_cast<uintptr_t>(p.ps.cols[c].ptr) & 15) == 0;
      for (int k = 0; ok && k < st->nkeys; k++) {
        const CompiledProgram& cpk = pb.prog(has_pred + k);
        ok = cpk.is_plain_column;
        sp.key_slot[k] = cpk.plain_slot;
      }
      for (int g = 0; ok && g < nargs; g++) {
        const CompiledProgram& cpa = pb.prog(has_pred + st->nkeys + g);
        ok = cpa.is_plain_column;
        sp.arg_slot[g] = cpa.plain_slot;
      }
      if (ok && has_pred) {
        // t0 [t1 AND|OR [t2 AND|OR ...]] in lowered form: (PUSH_COL, CMP leaf) {(PUSH_COL, CMP leaf), AND|OR stack}*
        const int b = p.ps.start[0], e = p.ps.start[1];
        const DevInsn* in = &p.ps.insn[b];
        auto term_at = [&](int i, PlainTerm* out) {
          if (i + 1 >= e - b) return false;
          const DevInsn &c = in[i], &o = in[i + 1];
          if (c.op != V_PUSH_COL) return false;
          if (o.op < V_EQ || o.op > V_GE || o.mode == RHS_STACK) return false;
          if (o.mode == RHS_COL && p.ps.cols[o.slot].dtype != p.ps.cols[c.slot].dtype) return false;
          memset(out, 0, sizeof(*out));
          out->kind = o.mode == RHS_COL ? 2 : 3;
          out->op = o.op;
          out->a = c.slot;
          out->b = o.mode == RHS_COL ? o.slot : 0;
          out->mt = mtype_of(p.ps.cols[c.slot].dtype);
          out->imm = o.imm;
          return true;
        };
        ok = term_at(0, &sp.term[0]);
        sp.nterms = ok ? 1 : 0;
        int i = 2;
        while (ok && i < e - b) {
          if (sp.nterms >= 4 || !term_at(i, &sp.term[sp.nterms]) || i + 2 >= e - b) { ok = false; break; }
          const DevInsn& j = in[i + 2];
          if ((j.op != V_AND && j.op != V_OR) || j.mode != RHS_STACK) { ok = false; break; }
          sp.term[sp.nterms].conn = j.op == V_OR ? 1 : 0;
          sp.nterms++;
          i += 3;
        }
      }
      sp.ncols = p.ps.ncols;
      if (ok) p.plain = sp;
      use_plain = ok;
    }
    // lean kernel: one 8-byte integer key column, one 8-byte argument column, distinct MIN/MAX/SUM/COUNT, no WHERE
    int lean_mask = 0, lean_mt = 0;
    {
      static const bool off = getenv("DFGPU_AGG_LEAN") && atoi(getenv("DFGPU_AGG_LEAN")) == 0;  // A/B switch
      auto w8 = [](int dt) { return dt == DFGPU_FLOAT64 || dt == DFGPU_INT64 || dt == DFGPU_UINT64; };
      bool ok = !off && use_plain && !has_pred && st->nkeys == 1 && nargs == 1 && w8(p.ps.cols[p.plain.key_slot[0]].dtype) &&
                w8(p.ps.cols[p.plain.arg_slot[0]].dtype);
      for (int a = 0; ok && a < st->naggs; a++) {
        const int bit = 1 << (st->descs[size_t(a)].func - 1);  // MIN 1, MAX 2, SUM 4, COUNT 8
        ok = !(lean_mask & bit);
        lean_mask |= bit;
      }
      if (ok) lean_mt = mtype_of(p.ps.cols[p.plain.arg_slot[0]].dtype);
      else lean_mask = 0;
    }

    // GROUP BY: run, then replay rows that could not get a slot after growing the table.
    // First big batch with no cardinality hint: a 1 Mi-row prefix is aggregated first; the number of
    // groups it produces decides the table layout (SoA while the hot sectors fit L2, AoS beyond) before
    // the bulk of the batch is touched.
    unsigned* ovf[2] = {(unsigned*)ctx->alloc(size_t(batch->nrows) * 4), nullptr};
    tr.mark("overflow list alloc");
    struct Freer {
      dfgpu_ctx* c;
      unsigned** o;
      ~Freer() { c->free(o[0]); c->free(o[1]); }
    } freer{ctx, ovf};
    const long long kPrefix = 1ll << 20;
    const bool sample = st->rows_seen == batch->nrows && st->expected == 0 && !st->aos && batch->nrows >= 4 * kPrefix;
    std::vector<std::pair<long long, long long>> ranges;  // (begin, count)
    if (sample) {
      ranges.push_back({0, kPrefix});
      ranges.push_back({kPrefix, batch->nrows - kPrefix});
    } else {
      ranges.push_back({0, batch->nrows});
    }
    for (size_t ri = 0; ri < ranges.size(); ri++) {
      int cur = 0;
      const unsigned* list = nullptr;
      long long nlist = 0;
      for (int round = 0;; round++) {
        if (round > 40) fail(DFGPU_ERR_INTERNAL, "hash table growth did not converge");
        p.t = st->t;
        p.cap = st->cap;
        p.max_groups = st->cap / 2;
        p.row_begin = ranges[ri].first;
        p.nrows = ranges[ri].second;
        p.row_list = list;
        p.nlist = nlist;
        p.ovf_rows = ovf[cur];
        DF_CUDA(cudaMemsetAsync(st->d_counters + 1, 0, 8, ctx->stream));
        const long long n = list ? nlist : p.nrows;
        const bool front = st->use_front && !list;
        // <= 64 groups: one private 256-slot table per warp (same shared-memory footprint as the
        // CTA-wide 2048-slot table); otherwise one table per CTA
        p.front_per_warp = st->ngroups <= 64 ? 1 : 0;
        p.front_slots = p.front_per_warp ? AG_FRONT_SLOTS / (AG_THREADS / 32) : AG_FRONT_SLOTS;
        {
          static const bool no_hint = getenv("DFGPU_AGG_STREAM_HINT") && atoi(getenv("DFGPU_AGG_STREAM_HINT")) == 0;  // A/B switch
          p.stream_hint = no_hint ? 0 : 1;
        }
        // (A persisting-L2 access-policy window over the table was measured in round 2 and removed: the scan
        // went from 1.57 to 7.8 ms at 1e5 groups and from 3.4 to 15 ms at 1e6, profiles/r02a_l2persist.txt.)
        // tables that outgrow L2: several passes over the rows, each confined to one contiguous part of the table
        const int npass = (list || front || ranges[ri].second < (4ll << 20) || st->aos) ? 1 : st->npass;
        p.npass = npass;
        p.pass_shift = 64;
        for (int q = 1; q < npass; q <<= 1) p.pass_shift--;
        const long long max_groups = p.max_groups;
        for (int pass = 0; pass < npass; pass++) {
          p.pass_id = pass;
          p.max_groups = max_groups;
          if (p.ps.has_nulls) launch_scan(ctx, k_hash_agg<8, false, true>, p, n, false);
          else if (lean_mask && !list && !front && !st->aos && (p.row_begin & 1) == 0) {
            // the lean kernel addresses the hybrid layout directly
            p.lean.key_col = (const unsigned long long*)p.ps.cols[p.plain.key_slot[0]].ptr;
            p.lean.arg_col = (const unsigned long long*)p.ps.cols[p.plain.arg_slot[0]].ptr;
            p.lean.min_w = p.lean.max_w = 0;
            p.lean.sum_arr = p.lean.cnt_arr = nullptr;
            for (int a = 0; a < st->naggs; a++) {
              const int f = st->descs[size_t(a)].func, l = st->t.loc[a];
              if (f == DFGPU_AGG_MIN) p.lean.min_w = l;
              else if (f == DFGPU_AGG_MAX) p.lean.max_w = l;
              else if (f == DFGPU_AGG_SUM) p.lean.sum_arr = st->t.add + (long long)(~l) * st->t.astride;
              else p.lean.cnt_arr = st->t.add + (long long)(~l) * st->t.astride;
            }
            bool layout_ok = st->t.lw == (((lean_mask & 3) == 3) ? 4 : ((lean_mask & 3) ? 2 : 1));
            for (int a = 0; a < st->naggs; a++) {
              const int f = st->descs[size_t(a)].func;
              layout_ok = layout_ok && ((f == DFGPU_AGG_MIN || f == DFGPU_AGG_MAX) ? st->t.loc[a] >= 1 : st->t.loc[a] < 0);
            }
            if (layout_ok) launch_lean(ctx, p, n, lean_mask, lean_mt);
            else launch_scan(ctx, k_hash_agg_plain<2, false, false>, p, n, false);
          } else if (use_plain && !list && (p.row_begin & 1) == 0) {
            static const bool pf = getenv("DFGPU_AGG_PREFETCH") && atoi(getenv("DFGPU_AGG_PREFETCH")) != 0;  // A/B switch (measured: no gain)
            const bool two = p.plain.ncols <= 2;
            if (front) {
              if (two) launch_scan(ctx, k_hash_agg_plain<2, false, true>, p, n, true);
              else launch_scan(ctx, k_hash_agg_plain<4, false, true>, p, n, true);
            } else if (!pf) {
              if (two) launch_scan(ctx, k_hash_agg_plain<2, false, false>, p, n, false);
              else launch_scan(ctx, k_hash_agg_plain<4, false, false>, p, n, false);
            } else {
              if (two) launch_scan(ctx, k_hash_agg_plain<2, true, false>, p, n, false);
              else launch_scan(ctx, k_hash_agg_plain<4, true, false>, p, n, false);
            }
          } else if (d <= 1) launch_hash_agg<1>(ctx, p, n, front);
          else if (d <= 2) launch_hash_agg<2>(ctx, p, n, front);
          else if (d <= 4) launch_hash_agg<4>(ctx, p, n, front);
          else launch_hash_agg<8>(ctx, p, n, front);
        }
        unsigned long long c[8];
        read_counters(st, c);
        if (c[3] == 2) fail(DFGPU_ERR_INTERNAL, "front-table merge could not find a slot");
        if (c[3]) fail(DFGPU_ERR_ARROW, "DivideByZero");
        st->ngroups = (long long)c[0];
        st->sentinel_used = c[2] != 0;
        const long long novf = (long long)c[1];
        tr.mark(ri == 0 && ranges.size() > 1 ? "scan kernel (prefix)" : "scan kernel");
        if (novf == 0) {
          if (st->ngroups > st->cap / 2) { table_grow(st, st->cap * 4); tr.mark("table_grow (load factor)"); }
          break;
        }
        table_grow(st, st->cap * 4);
        list = ovf[cur];
        nlist = novf;
        cur ^= 1;
        if (!ovf[cur]) ovf[cur] = (unsigned*)ctx->alloc(size_t(batch->nrows) * 4);
      }
      // few groups so far: later rows of this stream go through the shared-memory front table
      st->use_front = st->ngroups <= AG_FRONT_MAX_GROUPS && st->rows_seen >= (1ll << 20);
      if (sample && ri == 0) {
        // size (and lay out) the table for the estimated number of groups before the bulk of the batch
        // is touched: one rebuild of a ~1 Mi-entry table instead of repeated 4x growth + replays
        long long est = std::max(estimate_groups(st->ngroups, kPrefix, batch->nrows), st->ngroups);
        const long long afford = (long long)(ctx->device_mem_bytes / 8) / (32 * (1 + st->naggs));  // slots that fit in 1/8 of device memory
        long long want_cap = std::max(AG_MIN_CAP, next_pow2(est * 2));
        while (want_cap > st->cap && want_cap > afford) want_cap >>= 1;
        const bool to_aos = !st->aos && want_aos(est, st->descs, st->naggs);
        st->npass = (st->aos || to_aos) ? 1 : passes_for(est, st->descs, st->naggs);
        if (to_aos || want_cap > st->cap) {
          st->aos = st->aos || to_aos;
          table_grow(st, std::max(st->cap, want_cap));
          tr.mark("table_grow (prefix estimate)");
        }
      }
    }
    if (ukey) {
      VerifyParams vp;
      memset(&vp, 0, sizeof(vp));
      vp.hashes = d_hash;
      vp.n = batch->nrows;
      vp.t = st->t;
      vp.cap = st->cap;
      vp.rep_agg = st->naggs - 1;
      vp.off = ukey->offsets;
      vp.bytes = (const unsigned char*)ukey->values;
      vp.srcs = st->d_utf8_srcs;
      vp.flag = st->d_counters + 5;
      DF_CUDA(cudaMemsetAsync(st->d_counters + 5, 0, 8, ctx->stream));
      k_utf8_group_verify<<<grid_for(ctx, batch->nrows, 256, 8), 256, 0, ctx->stream>>>(vp);
      DF_CUDA(cudaGetLastError());
      ctx->launches++;
      DF_CUDA(cudaMemcpyAsync(ctx->h_scratch + 12, st->d_counters + 5, 8, cudaMemcpyDeviceToHost, ctx->stream));
      DF_CUDA(cudaStreamSynchronize(ctx->stream));
      if (ctx->h_scratch[12] == 1ull) fail(DFGPU_ERR_INTERNAL, "two different Utf8 GROUP BY keys share a 64-bit hash (p < 1e-7 per 1e6 distinct keys)");
      if (ctx->h_scratch[12]) fail(DFGPU_ERR_INTERNAL, "Utf8 GROUP BY verification could not find a group");
    }
  }
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// multi-GPU merge of partial aggregates.  Every rank ends with the global result.
//   no GROUP BY : one ncclAllReduce per accumulator (api.cu comm_allreduce_aggs)
//   GROUP BY    : open-addressed slots are not canonical across ranks, so the all-reduce is sparse and
//                 owner-partitioned:  raw-compact the local table -> count entries per owner rank ->
//                 all-gather a (header, counts) record -> scatter the entries into per-owner segments ->
//                 ONE grouped ncclSend/ncclRecv all-to-all -> each rank merges only the keys it owns into a
//                 fresh table (k_merge) -> compacts them -> all ranks gather the owned segments
//                 (grouped ncclBroadcast) and decode the same rows in the same order, so results are
//                 bit-identical on every rank.  Work per rank is O(G_local + G/W + G), not O(W x G).
// A rank that never saw a batch takes part with zero entries and adopts the key / argument types of a rank
// that did (they travel in the header).
// ---------------------------------------------------------------------------------------------
namespace {

void agg_export_raw(dfgpu_aggstate* st, unsigned long long** keys, unsigned long long** vals, long long* n) {
  dfgpu_ctx* ctx = st->ctx;
  const long long cnt = st->nkeys == 0 ? 1 : st->ngroups + (st->sentinel_used ? 1 : 0);
  const size_t alloc_n = size_t(cnt > 0 ? cnt : 1);
  *keys = (unsigned long long*)ctx->alloc(alloc_n * 8);
  *vals = (unsigned long long*)ctx->alloc(alloc_n * 8 * size_t(st->naggs));
  CompactParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.t = st->t;
  cp.cap = st->cap;
  cp.sentinel_used = st->nkeys == 0 ? 1 : (st->sentinel_used ? 1 : 0);
  cp.nkeys = st->nkeys;
  cp.naggs = st->naggs;
  cp.raw = 1;
  cp.out_keys[0] = *keys;
  for (int a = 0; a < st->naggs; a++) {
    cp.aggs[a] = st->descs[size_t(a)];
    cp.out_vals[a] = *vals + size_t(a) * alloc_n;
  }
  DF_CUDA(cudaMemsetAsync(st->d_counters + 4, 0, 8, ctx->stream));
  cp.counter = st->d_counters + 4;
  k_compact<<<grid_for(ctx, st->cap + 1, 256, 8), 256, 0, ctx->stream>>>(cp);
  DF_CUDA(cudaGetLastError());
  ctx->launches++;
  *n = cnt;
}

// no GROUP BY
void agg_exchange_scalars(dfgpu_ctx* ctx, dfgpu_aggstate* st) {
  int funcs[kMaxAggs], mtypes[kMaxAggs];
  for (int a = 0; a < st->naggs; a++) {
    funcs[a] = st->descs[size_t(a)].func;
    mtypes[a] = st->descs[size_t(a)].mtype;
  }
  // per-aggregate non-null input counts travel with the accumulators: fold the host-side counts of the
  // null-free batches into the device counters, which the exchange sums over ranks
  DF_CUDA(cudaMemcpyAsync(ctx->h_scratch + 40, st->d_counters + 8, 64, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int a = 0; a < kMaxAggs; a++) {
    if (a < st->naggs) ctx->h_scratch[40 + a] += (unsigned long long)st->nonnull_host[size_t(a)];
    if (a < st->naggs) st->nonnull_host[size_t(a)] = 0;
  }
  ctx->h_scratch[48] = (unsigned long long)st->rows_seen;
  DF_CUDA(cudaMemcpyAsync(st->d_counters + 8, ctx->h_scratch + 40, 64, cudaMemcpyHostToDevice, ctx->stream));
  DF_CUDA(cudaMemcpyAsync(st->d_counters + 7, ctx->h_scratch + 48, 8, cudaMemcpyHostToDevice, ctx->stream));
  st->saw_nulls = true;
  // slot 0's accumulators (cap = 0: every accumulator in its own one-word array, contiguous)
  comm_allreduce_aggs(ctx, st->naggs, funcs, mtypes, st->t.val(0, 0), st->d_counters + 8, st->d_counters + 7);
  DF_CUDA(cudaMemcpyAsync(ctx->h_scratch + 48, st->d_counters + 7, 8, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  st->rows_seen = (long long)ctx->h_scratch[48];
}

// GROUP BY.  Returns the global result as raw rows of (1 + naggs) words in *rows (caller frees) and their number.
constexpr int HDR = 16;  // header words: [0] typed [1] nkeys [2] naggs [3] utf8 [4] rows_seen [5] key dtypes (8 bits each) [6] arg dtypes (8 bits each)
struct GatheredSegments {
  int nseg = 0;
  long long stride = 0;
  long long n[AG_MAX_WORLD] = {0};
};
void agg_exchange_groups(dfgpu_ctx* ctx, dfgpu_aggstate* st, unsigned long long** rows_out, long long* n_out, GatheredSegments* seg, bool* regroup) {
  const int W = ctx->world, me = ctx->rank;
  if (W > AG_MAX_WORLD) fail(DFGPU_ERR_NOT_IMPLEMENTED, "more than " + std::to_string(AG_MAX_WORLD) + " ranks");
  Trace tr(ctx);
  std::vector<void*> owned;
  struct Freer { dfgpu_ctx* c; std::vector<void*>* v; ~Freer() { for (void* q : *v) c->free(q); } } freer{ctx, &owned};
  auto dalloc = [&](size_t words) { void* q = ctx->alloc((words ? words : 1) * 8); owned.push_back(q); return (unsigned long long*)q; };
  // 1. local entries, counted per owner
  unsigned long long *keys = nullptr, *vals = nullptr;
  long long n_local = 0;
  // Utf8 and wide keys do not fit the (packed key, accumulators) rows of this exchange: if any rank has them, every
  // rank leaves through the regroup merge (finish_regroup) right after the header round
  const bool special = st->typed && (st->wide || st->utf8_key);
  if (st->typed && !special) {
    agg_export_raw(st, &keys, &vals, &n_local);
    owned.push_back(keys);
    owned.push_back(vals);
  }
  unsigned long long* d_rec = dalloc(size_t(HDR + W));
  DF_CUDA(cudaMemsetAsync(d_rec, 0, size_t(HDR + W) * 8, ctx->stream));
  OwnerParams op;
  memset(&op, 0, sizeof(op));
  op.keys = keys;
  const size_t ln = size_t(n_local > 0 ? n_local : 1);
  for (int a = 0; a < st->naggs; a++) op.vals[a] = vals + size_t(a) * ln;
  op.n = n_local;
  op.world = W;
  op.naggs = st->naggs;
  op.counts = d_rec + HDR;
  if (n_local > 0) {
    k_owner_count<<<grid_for(ctx, n_local, 256 * 4, 8), 256, 0, ctx->stream>>>(op);
    DF_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  unsigned long long* hh = ctx->h_scratch + 16;  // pinned
  memset(hh, 0, HDR * 8);
  hh[0] = st->typed ? 1 : 0;
  hh[1] = (unsigned long long)st->nkeys;
  hh[2] = (unsigned long long)(st->utf8_key ? st->naggs - 1 : st->naggs);  // user aggregates
  hh[3] = special ? 1 : 0;
  hh[7] = st->utf8_key ? 1 : 0;  // the single-Utf8-key form carries a hidden representative aggregate
  hh[4] = (unsigned long long)st->rows_seen;
  if (st->typed) {
    for (int k = 0; k < st->nkeys; k++) hh[5] |= (unsigned long long)(st->key_dtypes[size_t(k)] & 0xff) << (8 * k);
    for (int a = 0; a < (st->utf8_key ? st->naggs - 1 : st->naggs); a++) hh[6] |= (unsigned long long)(st->descs[size_t(a)].dtype & 0xff) << (8 * a);
    if (st->utf8_key) hh[5] = DFGPU_UTF8;  // the key the caller sees (internally: a UInt64 string hash)
  }
  DF_CUDA(cudaMemcpyAsync(d_rec, hh, HDR * 8, cudaMemcpyHostToDevice, ctx->stream));
  // 2. all-gather (header, counts)
  unsigned long long* d_all = dalloc(size_t(W) * size_t(HDR + W));
  comm_allgather_u64(ctx, d_rec, d_all, size_t(HDR + W));
  std::vector<unsigned long long> all(size_t(W) * size_t(HDR + W));
  DF_CUDA(cudaMemcpyAsync(all.data(), d_all, all.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  auto hdr = [&](int r, int i) { return all[size_t(r) * size_t(HDR + W) + size_t(i)]; };
  auto cnt = [&](int from, int to) { return (size_t)all[size_t(from) * size_t(HDR + W) + size_t(HDR + to)]; };
  // 3. agree on types; adopt them on a rank that saw no batch
  int typed_rank = -1;
  long long total_rows = 0;
  for (int r = 0; r < W; r++) {
    total_rows += (long long)hdr(r, 4);
    if ((int)hdr(r, 1) != st->nkeys) fail(DFGPU_ERR_GENERAL, "ranks disagree on the GROUP BY expressions");
    if (hdr(r, 3)) *regroup = true;
    if (hdr(r, 0)) {
      if (typed_rank < 0) typed_rank = r;
      else if (hdr(r, 5) != hdr(typed_rank, 5) || hdr(r, 6) != hdr(typed_rank, 6) || hdr(r, 2) != hdr(typed_rank, 2) || hdr(r, 3) != hdr(typed_rank, 3))
        fail(DFGPU_ERR_GENERAL, "ranks disagree on GROUP BY key / aggregate argument types");
    }
  }
  st->rows_seen = total_rows;
  if (!st->typed) {
    st->key_dtypes.clear();
    st->descs.clear();
    for (int k = 0; k < st->nkeys; k++) st->key_dtypes.push_back(typed_rank >= 0 ? int((hdr(typed_rank, 5) >> (8 * k)) & 0xff) : DFGPU_INT64);
    for (int a = 0; a < st->naggs; a++) {
      int dt = typed_rank >= 0 ? int((hdr(typed_rank, 6) >> (8 * a)) & 0xff) : st->out_dtypes[size_t(a)];
      if (!is_numeric(dt)) dt = DFGPU_FLOAT64;
      AggDesc d;
      d.func = uint8_t(st->funcs[size_t(a)]);
      d.dtype = uint8_t(dt);
      d.mtype = mtype_of(dt);
      d.out_dtype = uint8_t(d.func == DFGPU_AGG_COUNT ? DFGPU_UINT64 : dt);
      st->descs.push_back(d);
    }
    st->key_shift.clear();
    st->key_mask.clear();
    int bits = 0;
    for (int k = st->nkeys - 1; k >= 0 && !*regroup; k--) {  // (the regroup merge never packs keys)
      const int w = dtype_width(st->key_dtypes[size_t(k)]) * 8;
      st->key_shift.insert(st->key_shift.begin(), bits);
      st->key_mask.insert(st->key_mask.begin(), w == 64 ? ~0ull : ((1ull << w) - 1ull));
      bits += w;
    }
    if (st->nkeys == 1 && !*regroup) st->key_mask[0] = ~0ull;
    st->typed = true;
  }
  if (*regroup) {
    *rows_out = nullptr;
    *n_out = -1;
    return;
  }
  const size_t E = size_t(1 + st->naggs);
  // 4. scatter the local entries into per-owner segments
  std::vector<size_t> s_off(size_t(W), 0), s_cnt(size_t(W), 0), r_off(size_t(W), 0), r_cnt(size_t(W), 0);
  size_t total_send = 0, total_recv = 0;
  for (int r = 0; r < W; r++) {
    s_off[size_t(r)] = total_send * E;
    s_cnt[size_t(r)] = cnt(me, r) * E;
    total_send += cnt(me, r);
    r_off[size_t(r)] = total_recv * E;
    r_cnt[size_t(r)] = cnt(r, me) * E;
    total_recv += cnt(r, me);
  }
  if ((long long)total_send != n_local) fail(DFGPU_ERR_INTERNAL, "owner counts do not add up to the local groups");
  size_t max_recv = 1;  // the most entries any rank receives (column sums of the count matrix)
  for (int r = 0; r < W; r++) {
    size_t col = 0;
    for (int q = 0; q < W; q++) col += cnt(q, r);
    max_recv = std::max(max_recv, col);
  }
  unsigned long long* d_send = dalloc(total_send * E);
  if (n_local > 0) {
    unsigned long long* d_cursor = dalloc(size_t(W));
    DF_CUDA(cudaMemsetAsync(d_cursor, 0, size_t(W) * 8, ctx->stream));
    op.cursor = d_cursor;
    op.rows = d_send;
    for (int r = 0; r < W; r++) op.seg_off[r] = s_off[size_t(r)] / E;
    k_owner_scatter<<<grid_for(ctx, n_local, 256 * 4, 8), 256, 0, ctx->stream>>>(op);
    DF_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  // 5. all-to-all: every entry goes to its owner
  unsigned long long* d_recv = dalloc(total_recv * E);
  comm_exchange_v(ctx, d_send, s_off.data(), s_cnt.data(), d_recv, r_off.data(), r_cnt.data());
  tr.mark("exchange: export + all-to-all");
  // 6. merge what this rank owns into a fresh table
  long long n_owned = 0;
  unsigned long long* d_owned = nullptr;
  if (total_recv > 0) {
    const long long ocap = std::max<long long>(1024, next_pow2(2 * (long long)total_recv));
    const bool oaos = want_aos((long long)total_recv, st->descs, st->naggs);
    TableLayout ot = table_alloc(ctx, st->naggs, st->nkeys, st->descs, ocap, oaos);
    owned.push_back(ot.base);
    MergeParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.in_keys = d_recv;
    for (int a = 0; a < st->naggs; a++) {
      mp.in_vals[a] = d_recv + 1 + a;
      mp.aggs[a] = st->descs[size_t(a)];
    }
    mp.in_stride = (long long)E;
    mp.n = (long long)total_recv;
    mp.t = ot;
    mp.cap = ocap;
    mp.naggs = st->naggs;
    DF_CUDA(cudaMemsetAsync(st->d_counters, 0, 32, ctx->stream));
    mp.counters = st->d_counters;
    k_merge<<<grid_for(ctx, mp.n, 256, 8), 256, 0, ctx->stream>>>(mp);
    DF_CUDA(cudaGetLastError());
    ctx->launches++;
    // compact what this rank owns without a host round trip in between: the buffer has room for the largest
    // number of entries any rank receives (known to all from the count matrix), which also makes it a valid send
    // buffer of the padded all-gather below; the sentinel slot's use and the count stay on the device
    d_owned = dalloc(max_recv * E);
    CompactParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.t = ot;
    cp.cap = ocap;
    cp.sentinel_used = 0;
    cp.sentinel_flag = st->d_counters + 2;
    cp.nkeys = st->nkeys;
    cp.naggs = st->naggs;
    cp.raw = 1;
    cp.raw_stride = (long long)E;
    cp.out_keys[0] = d_owned;
    for (int a = 0; a < st->naggs; a++) {
      cp.aggs[a] = st->descs[size_t(a)];
      cp.out_vals[a] = d_owned + 1 + a;
    }
    DF_CUDA(cudaMemsetAsync(st->d_counters + 4, 0, 8, ctx->stream));
    cp.counter = st->d_counters + 4;
    k_compact<<<grid_for(ctx, ocap + 1, 256, 8), 256, 0, ctx->stream>>>(cp);
    DF_CUDA(cudaGetLastError());
    ctx->launches++;
  } else {
    DF_CUDA(cudaMemsetAsync(st->d_counters, 0, 64, ctx->stream));
  }
  // 7. every rank gathers the owned segments: sizes first ([owned entries, merge error flag] per rank)
  unsigned long long* d_n = dalloc(size_t(2 + 2 * W));
  DF_CUDA(cudaMemcpyAsync(d_n, st->d_counters + 4, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  DF_CUDA(cudaMemcpyAsync(d_n + 1, st->d_counters + 3, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  comm_allgather_u64(ctx, d_n, d_n + 2, 2);
  std::vector<unsigned long long> owned_n2(size_t(2 * W), 0);
  DF_CUDA(cudaMemcpyAsync(owned_n2.data(), d_n + 2, size_t(2 * W) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<unsigned long long> owned_n(size_t(W), 0);
  for (int r = 0; r < W; r++) {
    if (owned_n2[size_t(2 * r + 1)]) fail(DFGPU_ERR_INTERNAL, "partial-aggregate merge failed on rank " + std::to_string(r));
    owned_n[size_t(r)] = owned_n2[size_t(2 * r)];
  }
  n_owned = (long long)owned_n[size_t(me)];
  // one fixed-size all-gather of max_owned entries per rank (owners are a hash of the key, so the segments are
  // within a few percent of each other): one collective instead of one broadcast per rank
  size_t max_owned = 0, G = 0;
  for (int r = 0; r < W; r++) {
    max_owned = std::max(max_owned, size_t(owned_n[size_t(r)]));
    G += size_t(owned_n[size_t(r)]);
  }
  if (max_owned > max_recv) fail(DFGPU_ERR_INTERNAL, "owned groups exceed the received entries");
  unsigned long long* d_final = (unsigned long long*)ctx->alloc((max_owned ? size_t(W) * max_owned * E : 1) * 8);
  if (max_owned > 0) {
    if (!d_owned) d_owned = dalloc(max_recv * E);  // a rank that owns nothing still contributes its (empty) segment
    comm_allgather_u64(ctx, d_owned, d_final, max_owned * E);
  }
  DF_CUDA(cudaStreamSynchronize(ctx->stream));  // the temporaries above are released on return
  tr.mark("exchange: owner merge + gather");
  *rows_out = d_final;
  *n_out = (long long)G;
  seg->nseg = W;
  seg->stride = (long long)max_owned;
  for (int r = 0; r < W; r++) seg->n[r] = (long long)owned_n[size_t(r)];
}

}  // namespace

// Multi-GPU merge for key shapes whose groups cannot travel as (packed key, accumulators) rows — Utf8 keys and
// wide composite keys: every rank finishes locally, the ranks all-gather their LOCAL RESULT columns (strings
// included), and every rank aggregates the concatenation once more with the merge function of each aggregate
// (SUM -> SUM, COUNT -> SUM of counts, MIN -> MIN, MAX -> MAX) through the same single-GPU operator.  O(W x G) work
// per rank instead of the owner-partitioned O(G): accepted for these shapes.  Results agree across ranks bit for
// bit except Float64 SUMs (order of the second aggregation's reductions: <= 1e-9 relative).
namespace {
struct WorldGuard {  // run a stretch of the operator as if no communicator were attached
  dfgpu_ctx* c;
  int world;
  explicit WorldGuard(dfgpu_ctx* ctx) : c(ctx), world(ctx->world) { c->world = 1; }
  ~WorldGuard() { c->world = world; }
};

std::unique_ptr<dfgpu_result> finish_regroup(dfgpu_ctx* ctx, dfgpu_aggstate* st) {
  const int W = ctx->world, me = ctx->rank;
  const int user_aggs = st->utf8_key ? st->naggs - 1 : st->naggs;
  const int ncols = st->nkeys + user_aggs;
  // 1. this rank's own result (an empty one when it saw no batch: types were adopted from the header)
  std::unique_ptr<dfgpu_result> local;
  if (st->t.base) {
    WorldGuard g(ctx);
    dfgpu_result* r = nullptr;
    const int rc = dfgpu_aggregate_finish(st, &r);
    if (rc != DFGPU_OK) fail(rc, dfgpu_last_error());
    local.reset(r);
  } else {
    local = std::make_unique<dfgpu_result>();
    local->ctx = ctx;
    local->nrows = 0;
    for (int k = 0; k < st->nkeys; k++) {
      DevColumn c;
      c.dtype = st->key_dtypes[size_t(k)];
      if (c.dtype == DFGPU_UTF8) {
        c.offsets = (int32_t*)ctx->alloc(4);
        DF_CUDA(cudaMemsetAsync(c.offsets, 0, 4, ctx->stream));
      }
      local->cols.push_back(c);
    }
    for (int a = 0; a < user_aggs; a++) {
      DevColumn c;
      c.dtype = st->descs[size_t(a)].out_dtype;
      local->cols.push_back(c);
    }
  }
  if (int(local->cols.size()) != ncols) fail(DFGPU_ERR_INTERNAL, "regroup merge: unexpected local result shape");
  // 2. sizes: [rows, bytes of every Utf8 column] per rank
  const int NH = 1 + kMaxKeys;
  std::vector<void*> tmp;
  struct Freer { dfgpu_ctx* c; std::vector<void*>* v; ~Freer() { for (void* q : *v) c->free(q); } } freer{ctx, &tmp};
  auto dalloc = [&](size_t bytes) { void* q = ctx->alloc(bytes ? bytes : 8); tmp.push_back(q); return q; };
  unsigned long long* d_h = (unsigned long long*)dalloc(size_t(NH) * 8 * size_t(W + 1));
  unsigned long long* hh = ctx->h_scratch + 16;
  memset(hh, 0, size_t(NH) * 8);
  hh[0] = (unsigned long long)local->nrows;
  for (int k = 0; k < st->nkeys; k++)
    if (local->cols[size_t(k)].dtype == DFGPU_UTF8) hh[1 + k] = (unsigned long long)local->cols[size_t(k)].values_bytes;
  DF_CUDA(cudaMemcpyAsync(d_h, hh, size_t(NH) * 8, cudaMemcpyHostToDevice, ctx->stream));
  comm_allgather_u64(ctx, d_h, d_h + NH, size_t(NH));
  std::vector<unsigned long long> all(size_t(NH) * size_t(W));
  DF_CUDA(cudaMemcpyAsync(all.data(), d_h + NH, all.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  long long N = 0;
  std::vector<long long> row_base(size_t(W), 0);
  for (int r = 0; r < W; r++) {
    row_base[size_t(r)] = N;
    N += (long long)all[size_t(r) * NH];
  }
  if (N >= (1ll << 31)) fail(DFGPU_ERR_NOT_IMPLEMENTED, "regroup merge of 2^31 or more partial groups");
  // 3. gather every result column
  auto gathered = std::make_unique<dfgpu_batch>();
  gathered->ctx = ctx;
  gathered->nrows = N;
  std::vector<size_t> off(size_t(W), 0), cnt(size_t(W), 0);
  for (int c = 0; c < ncols; c++) {
    const DevColumn& lc = local->cols[size_t(c)];
    DevColumn gc;
    gc.dtype = lc.dtype;
    if (lc.dtype == DFGPU_UTF8) {
      // bytes
      size_t total_b = 0;
      std::vector<size_t> bbase(size_t(W), 0);
      for (int r = 0; r < W; r++) {
        bbase[size_t(r)] = total_b;
        off[size_t(r)] = total_b;
        cnt[size_t(r)] = size_t(all[size_t(r) * NH + 1 + size_t(c)]);
        total_b += cnt[size_t(r)];
      }
      if (total_b >= (1ull << 31)) fail(DFGPU_ERR_NOT_IMPLEMENTED, "regroup merge of 2 GiB or more of key strings");
      gc.values_bytes = total_b;
      gc.values = ctx->alloc(total_b ? total_b : 1);
      comm_allgather_bytes_v(ctx, lc.values, gc.values, off.data(), cnt.data());
      // offsets: every rank's (rows + 1) array, spliced with its byte base
      int* raw = (int*)dalloc(size_t(N + W) * 4);
      for (int r = 0; r < W; r++) {
        off[size_t(r)] = size_t(row_base[size_t(r)] + r) * 4;
        cnt[size_t(r)] = size_t(all[size_t(r) * NH] + 1) * 4;
      }
      comm_allgather_bytes_v(ctx, lc.offsets, raw, off.data(), cnt.data());
      gc.offsets = (int32_t*)ctx->alloc(size_t(N + 1) * 4);
      for (int r = 0; r < W; r++)
        shift_copy_i32(ctx, gc.offsets + row_base[size_t(r)], raw + row_base[size_t(r)] + r, (long long)all[size_t(r) * NH], int(bbase[size_t(r)]));
      const int last = int(total_b);
      DF_CUDA(cudaMemcpyAsync(gc.offsets + N, &last, 4, cudaMemcpyHostToDevice, ctx->stream));
      DF_CUDA(cudaStreamSynchronize(ctx->stream));  // `last` is a stack variable
    } else {
      const size_t w = size_t(dtype_width(lc.dtype));
      for (int r = 0; r < W; r++) {
        off[size_t(r)] = size_t(row_base[size_t(r)]) * w;
        cnt[size_t(r)] = size_t(all[size_t(r) * NH]) * w;
      }
      gc.values_bytes = size_t(N) * w;
      gc.values = ctx->alloc(gc.values_bytes ? gc.values_bytes : 8);
      comm_allgather_bytes_v(ctx, lc.values, gc.values, off.data(), cnt.data());
    }
    gathered->cols.push_back(gc);
  }
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  (void)me;
  // 4. aggregate the partial results once more, with each aggregate's merge function
  const size_t nk = size_t(st->nkeys), na = size_t(user_aggs);
  std::vector<dfgpu_insn> kprog(nk), aprog(na);
  std::vector<const dfgpu_insn*> kptr;
  std::vector<int> klen;
  for (int k = 0; k < st->nkeys; k++) {
    memset(&kprog[size_t(k)], 0, sizeof(dfgpu_insn));
    kprog[size_t(k)].op = DFGPU_OP_COL;
    kprog[size_t(k)].col = k;
    kprog[size_t(k)].dtype = gathered->cols[size_t(k)].dtype;
    kptr.push_back(&kprog[size_t(k)]);
    klen.push_back(1);
  }
  std::vector<dfgpu_agg> aggs(na);
  for (int a = 0; a < user_aggs; a++) {
    memset(&aprog[size_t(a)], 0, sizeof(dfgpu_insn));
    aprog[size_t(a)].op = DFGPU_OP_COL;
    aprog[size_t(a)].col = st->nkeys + a;
    aprog[size_t(a)].dtype = gathered->cols[size_t(st->nkeys + a)].dtype;
    const int f = st->descs[size_t(a)].func;
    aggs[size_t(a)].func = f == DFGPU_AGG_COUNT ? DFGPU_AGG_SUM : f;
    aggs[size_t(a)].arg = &aprog[size_t(a)];
    aggs[size_t(a)].arg_len = 1;
    aggs[size_t(a)].out_dtype = gathered->cols[size_t(st->nkeys + a)].dtype;
    aggs[size_t(a)]._pad = 0;
  }
  WorldGuard g(ctx);
  dfgpu_aggstate* st2 = nullptr;
  int rc = dfgpu_aggregate_create(ctx, kptr.data(), klen.data(), st->nkeys, aggs.data(), user_aggs, N / W + 1, &st2);
  if (rc != DFGPU_OK) fail(rc, dfgpu_last_error());
  struct StFree { dfgpu_aggstate* s; ~StFree() { dfgpu_aggregate_free(s); } } stfree{st2};
  dfgpu_result* res = nullptr;
  if (N > 0) {
    rc = dfgpu_aggregate_update(st2, gathered.get());
    if (rc != DFGPU_OK) fail(rc, dfgpu_last_error());
    rc = dfgpu_aggregate_finish(st2, &res);
    if (rc != DFGPU_OK) fail(rc, dfgpu_last_error());
    return std::unique_ptr<dfgpu_result>(res);
  }
  return local;  // nobody had a group: the (empty) local result has the right columns
}
}  // namespace

extern "C" int dfgpu_aggregate_finish(dfgpu_aggstate* st, dfgpu_result** out) {
  return guarded([&] {
    if (!st || !out) fail(DFGPU_ERR_GENERAL, "dfgpu_aggregate_finish: null argument");
    if (st->finished) fail(DFGPU_ERR_GENERAL, "aggregate already finished");  // one-shot (aggregate.rs:616-619)
    dfgpu_ctx* ctx = st->ctx;
    ctx->use();
    Trace tr(ctx);
    if (!st->typed && !(st->nkeys > 0 && ctx->world > 1)) {
      // no batch was ever seen: resolve types from the declared output types
      if (st->nkeys > 0) {
        // an empty GROUP BY input yields an empty batch; key types are unknown -> need a batch
        fail(DFGPU_ERR_GENERAL, "aggregate finished before any input batch was provided");
      }
      for (int a = 0; a < st->naggs; a++) {
        AggDesc d;
        int odt = st->out_dtypes[size_t(a)];
        if (!is_numeric(odt)) fail(DFGPU_ERR_GENERAL, "aggregate output type must be given when there is no input");
        d.func = uint8_t(st->funcs[size_t(a)]);
        d.dtype = uint8_t(odt);
        d.mtype = mtype_of(odt);
        d.out_dtype = uint8_t(odt);
        st->descs.push_back(d);
      }
      st->typed = true;
      st->cap = 0;
      st->t = table_alloc(ctx, st->naggs, st->nkeys, st->descs, 0, false);
    }
    // multi-GPU: every rank must take part (also one that saw no batch), and every rank gets the global result
    unsigned long long* xrows = nullptr;
    long long xn = -1;
    GatheredSegments xseg;
    bool regroup = false;
    struct XFree { dfgpu_ctx* c; unsigned long long** p; ~XFree() { c->free(*p); } } xfree{ctx, &xrows};
    if (ctx->world > 1) {
      if (st->nkeys == 0) agg_exchange_scalars(ctx, st);
      else agg_exchange_groups(ctx, st, &xrows, &xn, &xseg, &regroup);
      if (regroup) {
        *out = finish_regroup(ctx, st).release();
        st->finished = true;
        return;
      }
    }
    auto res = std::make_unique<dfgpu_result>();
    res->ctx = ctx;
    const long long cnt = xn >= 0 ? xn : (st->nkeys == 0 ? 1 : st->ngroups + (st->sentinel_used ? 1 : 0));
    const size_t alloc_n = size_t(cnt > 0 ? cnt : 1);
    CompactParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.t = st->t;
    cp.cap = st->cap;
    cp.sentinel_used = st->nkeys == 0 ? 1 : (st->sentinel_used ? 1 : 0);
    cp.nkeys = st->nkeys;
    cp.naggs = st->naggs;
    cp.raw = 0;
    cp.wide_kw = st->wide ? st->nkeys : 0;
    for (int k = 0; k < st->nkeys && st->wide; k++) cp.key_is_utf8[k] = st->key_is_utf8[size_t(k)];
    dfgpu_result hidden;  // RAII: hash-key and representative columns of a Utf8-keyed aggregate
    hidden.ctx = ctx;
    std::vector<std::pair<int, size_t>> wide_refs;  // wide keys: (key part, index of its reference column in `hidden`)
    for (int k = 0; k < st->nkeys; k++) {  // group columns first (aggregate.rs:890-925)
      DevColumn c;
      c.dtype = st->key_dtypes[size_t(k)];
      if (st->wide && st->key_is_utf8[size_t(k)]) {
        // references to the groups' strings come out of the compaction; the strings are gathered below
        c.dtype = DFGPU_UINT64;
        c.values_bytes = alloc_n * 8;
        c.values = ctx->alloc(c.values_bytes);
        wide_refs.push_back({k, hidden.cols.size()});
        hidden.cols.push_back(c);
        DevColumn u;
        u.dtype = DFGPU_UTF8;
        res->cols.push_back(u);
        cp.out_keys[k] = c.values;
        cp.key_dtype[k] = DFGPU_UINT64;
        continue;
      }
      c.values_bytes = alloc_n * size_t(dtype_width(c.dtype));
      c.values = ctx->alloc(c.values_bytes);
      if (st->utf8_key) {
        hidden.cols.push_back(c);
        DevColumn u;
        u.dtype = DFGPU_UTF8;  // filled by the gather below
        res->cols.push_back(u);
      } else {
        res->cols.push_back(c);
      }
      cp.out_keys[k] = c.values;
      cp.key_dtype[k] = c.dtype;
      cp.key_mask[k] = st->key_mask[size_t(k)];
      cp.key_shift[k] = st->key_shift[size_t(k)];
    }
    for (int a = 0; a < st->naggs; a++) {  // then aggregate columns (aggregate.rs:928-949)
      DevColumn c;
      c.dtype = st->descs[size_t(a)].out_dtype;
      c.values_bytes = alloc_n * size_t(dtype_width(c.dtype));
      c.values = ctx->alloc(c.values_bytes);
      if (st->utf8_key && a == st->naggs - 1) hidden.cols.push_back(c);  // representative rows: not a result column
      else res->cols.push_back(c);
      cp.out_vals[a] = c.values;
      cp.aggs[a] = st->descs[size_t(a)];
    }
    if (xn >= 0) {
      // the merged global rows came back from the exchange: decode them (same rows, same order on every rank)
      DecodeParams dp;
      memset(&dp, 0, sizeof(dp));
      dp.rows = xrows;
      dp.n = xn;
      dp.nseg = xseg.nseg;
      dp.seg_stride = xseg.stride;
      for (int r = 0; r < xseg.nseg; r++) dp.seg_n[r] = xseg.n[r];
      dp.nkeys = st->nkeys;
      dp.naggs = st->naggs;
      for (int k = 0; k < st->nkeys; k++) {
        dp.key_mask[k] = cp.key_mask[k];
        dp.key_shift[k] = cp.key_shift[k];
        dp.key_dtype[k] = cp.key_dtype[k];
        dp.out_keys[k] = cp.out_keys[k];
      }
      for (int a = 0; a < st->naggs; a++) {
        dp.aggs[a] = cp.aggs[a];
        dp.out_vals[a] = cp.out_vals[a];
      }
      if (xn > 0) {
        k_decode_rows<<<grid_for(ctx, xn, 256, 8), 256, 0, ctx->stream>>>(dp);
        DF_CUDA(cudaGetLastError());
        ctx->launches++;
      }
      DF_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
      DF_CUDA(cudaMemsetAsync(st->d_counters + 4, 0, 8, ctx->stream));
      cp.counter = st->d_counters + 4;
      k_compact<<<grid_for(ctx, st->cap + 1, 256, 8), 256, 0, ctx->stream>>>(cp);
      DF_CUDA(cudaGetLastError());
      ctx->launches++;
      DF_CUDA(cudaMemcpyAsync(ctx->h_scratch + 8, st->d_counters + 4, 8, cudaMemcpyDeviceToHost, ctx->stream));
      DF_CUDA(cudaStreamSynchronize(ctx->stream));
      if ((long long)ctx->h_scratch[8] != cnt) fail(DFGPU_ERR_INTERNAL, "table compaction count mismatch");
    }
    res->nrows = cnt;
    for (auto& wr : wide_refs)  // wide keys: the Utf8 parts' strings, in output order
      gather_utf8_multi(ctx, st->d_utf8_srcs, (const unsigned long long*)hidden.cols[wr.second].values, cnt, &res->cols[size_t(wr.first)]);
    if (st->utf8_key)  // key strings = the representatives' strings, in output order
      gather_utf8_multi(ctx, st->d_utf8_srcs, (const unsigned long long*)hidden.cols[1].values, cnt, &res->cols[0]);
    if (st->nkeys == 0) {
      // an aggregate that saw no non-null input is null (array_from_scalar!, aggregate.rs:641-643)
      std::vector<long long> nonnull = st->nonnull_host;
      if (st->saw_nulls) {
        DF_CUDA(cudaMemcpyAsync(ctx->h_scratch + 40, st->d_counters + 8, 64, cudaMemcpyDeviceToHost, ctx->stream));
        DF_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int a = 0; a < st->naggs; a++) nonnull[size_t(a)] += (long long)ctx->h_scratch[40 + a];
      }
      for (int a = 0; a < st->naggs; a++) {
        if (nonnull[size_t(a)] > 0 || (st->rows_seen > 0 && st->descs[size_t(a)].func == DFGPU_AGG_COUNT)) continue;
        DevColumn& c = res->cols[size_t(a)];
        c.validity = (uint8_t*)ctx->alloc(1);
        DF_CUDA(cudaMemsetAsync(c.validity, 0, 1, ctx->stream));
        c.null_count = 1;
      }
      DF_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    tr.mark("finish (compact + outputs)");
    st->finished = true;
    *out = res.release();
  });
}

extern "C" int dfgpu_aggregate_free(dfgpu_aggstate* st) {
  return guarded([&] {
    if (st) st->ctx->use();
    delete st;
  });
}
