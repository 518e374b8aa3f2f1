// expr_compile.cu — host half of the expression VM: type-check a postfix program the way
// compile_scalar_expr would (src/execution/expression.rs:283-505), fold right-hand leaves into the
// consuming instruction, and emit the device bytecode of expr_vm.cuh.
#include <memory>

#include "expr_vm.cuh"

namespace dfgpu {

MType mtype_of(int dt) {
  switch (dt) {
    case DFGPU_FLOAT64: return MT_F64;
    case DFGPU_FLOAT32: return MT_F32;
    case DFGPU_BOOL: return MT_BOOL;
    case DFGPU_INT8: case DFGPU_INT16: case DFGPU_INT32: case DFGPU_INT64: return MT_I;
    case DFGPU_UINT8: case DFGPU_UINT16: case DFGPU_UINT32: case DFGPU_UINT64: return MT_U;
  }
  return MT_NONE;
}

namespace {

struct Node {
  enum Kind { COL, LIT, CAST, BIN } kind = COL;
  int col = 0;
  int dtype = 0;             // result dtype
  unsigned long long imm = 0;  // LIT payload, widened to the machine representation
  int op = 0;                // DFGPU_OP_* for BIN
  std::unique_ptr<Node> l, r;
};

const char* op_debug_name(int op) {
  switch (op) {
    case DFGPU_OP_ADD: return "Plus"; case DFGPU_OP_SUB: return "Minus"; case DFGPU_OP_MUL: return "Multiply";
    case DFGPU_OP_DIV: return "Divide"; case DFGPU_OP_EQ: return "Eq"; case DFGPU_OP_NE: return "NotEq";
    case DFGPU_OP_LT: return "Lt"; case DFGPU_OP_LE: return "LtEq"; case DFGPU_OP_GT: return "Gt";
    case DFGPU_OP_GE: return "GtEq"; case DFGPU_OP_AND: return "And"; case DFGPU_OP_OR: return "Or";
  }
  return "?";
}

unsigned long long widen_literal(const dfgpu_insn& in) {
  switch (in.dtype) {
    case DFGPU_FLOAT64: case DFGPU_INT64: case DFGPU_UINT64: return in.lit.u64;
    case DFGPU_FLOAT32: { uint32_t b; memcpy(&b, &in.lit.f32, 4); return b; }
    case DFGPU_INT32: return (unsigned long long)(long long)(int32_t)in.lit.i64;
    case DFGPU_INT16: return (unsigned long long)(long long)(int16_t)in.lit.i64;
    case DFGPU_INT8: return (unsigned long long)(long long)(int8_t)in.lit.i64;
    case DFGPU_UINT32: return in.lit.u64 & 0xffffffffull;
    case DFGPU_UINT16: return in.lit.u64 & 0xffffull;
    case DFGPU_UINT8: return in.lit.u64 & 0xffull;
  }
  return 0;
}

VOp vop_of(int op) {
  switch (op) {
    case DFGPU_OP_ADD: return V_ADD; case DFGPU_OP_SUB: return V_SUB; case DFGPU_OP_MUL: return V_MUL;
    case DFGPU_OP_DIV: return V_DIV; case DFGPU_OP_EQ: return V_EQ; case DFGPU_OP_NE: return V_NE;
    case DFGPU_OP_LT: return V_LT; case DFGPU_OP_LE: return V_LE; case DFGPU_OP_GT: return V_GT;
    case DFGPU_OP_GE: return V_GE; case DFGPU_OP_AND: return V_AND; default: return V_OR;
  }
}

}  // namespace

int ProgramBuilder::slot_of_column(int col) {
  for (size_t i = 0; i < slots_.size(); i++)
    if (slots_[i] == col) return int(i);
  if (int(slots_.size()) >= kMaxCols)
    fail(DFGPU_ERR_NOT_IMPLEMENTED, "expression set references more than " + std::to_string(kMaxCols) + " distinct columns");
  slots_.push_back(col);
  return int(slots_.size()) - 1;
}

int ProgramBuilder::add(const dfgpu_insn* p, int n, const char* what) {
  if (n <= 0 || !p) fail(DFGPU_ERR_GENERAL, std::string("empty expression program for ") + what);
  // 1. postfix -> tree, with the reference's type rules
  std::vector<std::unique_ptr<Node>> st;
  for (int i = 0; i < n; i++) {
    auto nd = std::make_unique<Node>();
    const dfgpu_insn& in = p[i];
    switch (in.op) {
      case DFGPU_OP_COL: {  // Expr::Column (expression.rs:311-315)
        if (in.col < 0 || size_t(in.col) >= batch_->cols.size())
          fail(DFGPU_ERR_INVALID_COLUMN, "column index " + std::to_string(in.col) + " out of range");
        nd->kind = Node::COL;
        nd->col = in.col;
        nd->dtype = batch_->cols[size_t(in.col)].dtype;
        break;
      }
      case DFGPU_OP_LIT: {  // Expr::Literal (expression.rs:289-310)
        if (!is_numeric(in.dtype))
          fail(DFGPU_ERR_EXECUTION, std::string("No support for literal type ") + dtype_name(in.dtype));
        nd->kind = Node::LIT;
        nd->dtype = in.dtype;
        nd->imm = widen_literal(in);
        break;
      }
      case DFGPU_OP_CAST: {  // Expr::Cast (expression.rs:316-378)
        if (st.empty()) fail(DFGPU_ERR_GENERAL, "malformed expression program");
        auto inner = std::move(st.back());
        st.pop_back();
        if (inner->kind == Node::LIT) {
          // only Literal Int64 -> Float64 exists in the reference (expression.rs:345-373)
          if (inner->dtype != DFGPU_INT64)
            fail(DFGPU_ERR_NOT_IMPLEMENTED, std::string("CAST from ") + dtype_name(inner->dtype) + " to " + dtype_name(in.dtype));
          if (in.dtype != DFGPU_FLOAT64)
            fail(DFGPU_ERR_NOT_IMPLEMENTED, std::string("CAST from Int64 to ") + dtype_name(in.dtype));
          double d = double((long long)inner->imm);
          nd->kind = Node::LIT;
          nd->dtype = DFGPU_FLOAT64;
          memcpy(&nd->imm, &d, 8);
        } else if (inner->kind == Node::COL) {
          // The reference casts columns to Int16/Int32 only and panics otherwise
          // (cast_column_outer!, expression.rs:272-280).  Any numeric -> numeric cast is done here
          // (needed for the planner's own CAST(#i AS Int64) output, sqlplanner.rs:581); the result
          // type reported is the TARGET type (the reference reports the source: expression.rs:324).
          if (!is_numeric(inner->dtype) || !is_numeric(in.dtype))
            fail(DFGPU_ERR_NOT_IMPLEMENTED, std::string("CAST column from ") + dtype_name(inner->dtype) + " to " + dtype_name(in.dtype));
          nd->kind = Node::CAST;
          nd->dtype = in.dtype;
          nd->l = std::move(inner);
        } else {
          fail(DFGPU_ERR_GENERAL, "CAST not implemented for expression");  // expression.rs:374-377
        }
        break;
      }
      default: {
        int op = in.op;
        bool is_math = op >= DFGPU_OP_ADD && op <= DFGPU_OP_DIV;
        bool is_cmp = op >= DFGPU_OP_EQ && op <= DFGPU_OP_GE;
        bool is_bool = op == DFGPU_OP_AND || op == DFGPU_OP_OR;
        if (!is_math && !is_cmp && !is_bool) fail(DFGPU_ERR_EXECUTION, "operator: " + std::to_string(op));
        if (st.size() < 2) fail(DFGPU_ERR_GENERAL, "malformed expression program");
        nd->kind = Node::BIN;
        nd->op = op;
        nd->r = std::move(st.back());
        st.pop_back();
        nd->l = std::move(st.back());
        st.pop_back();
        int lt = nd->l->dtype, rt = nd->r->dtype;
        if (is_bool) {
          if (lt != DFGPU_BOOL || rt != DFGPU_BOOL)
            fail(DFGPU_ERR_INTERNAL, "boolean_ops: operand is not a BooleanArray (the reference panics here: expression.rs:217-221)");
          nd->dtype = DFGPU_BOOL;
        } else {
          if (lt != rt || !is_numeric(lt)) fail(DFGPU_ERR_EXECUTION, is_cmp ? "comparison_ops" : "math_ops");
          nd->dtype = is_cmp ? DFGPU_BOOL : lt;
        }
        break;
      }
    }
    st.push_back(std::move(nd));
  }
  if (st.size() != 1) fail(DFGPU_ERR_GENERAL, "malformed expression program");

  // 2. tree -> bytecode with right-hand leaf folding; track the live register-stack depth
  CompiledProgram cp;
  {
    // can the result be null?  (columns with nulls propagate through arithmetic / And / Or / Cast;
    // comparisons never produce nulls)
    struct N {
      const dfgpu_batch* b;
      bool go(const Node* nd) const {
        switch (nd->kind) {
          case Node::COL: return b->cols[size_t(nd->col)].null_count > 0;
          case Node::LIT: return false;
          case Node::CAST: return go(nd->l.get());
          default: {
            const bool cmp = nd->op >= DFGPU_OP_EQ && nd->op <= DFGPU_OP_GE;
            return !cmp && (go(nd->l.get()) || go(nd->r.get()));
          }
        }
      }
    } nn{batch_};
    cp.nullable = nn.go(st[0].get());
  }
  int depth = 0;
  struct Emit {
    ProgramBuilder* pb;
    CompiledProgram* cp;
    int* depth;
    void bump(int d) {
      *depth += d;
      if (*depth > cp->max_depth) cp->max_depth = *depth;
    }
    void go(const Node* nd) {
      DevInsn di;
      memset(&di, 0, sizeof(di));
      switch (nd->kind) {
        case Node::COL:
          di.op = V_PUSH_COL;
          di.slot = int16_t(pb->slot_of_column(nd->col));
          di.dtype = uint8_t(nd->dtype);
          di.mtype = mtype_of(nd->dtype);
          cp->code.push_back(di);
          bump(1);
          break;
        case Node::LIT:
          di.op = V_PUSH_IMM;
          di.imm = nd->imm;
          di.dtype = uint8_t(nd->dtype);
          di.mtype = mtype_of(nd->dtype);
          cp->code.push_back(di);
          bump(1);
          break;
        case Node::CAST:
          go(nd->l.get());
          di.op = V_CAST;
          di.dtype = uint8_t(nd->dtype);
          di.aux = int16_t(nd->l->dtype);
          di.mtype = mtype_of(nd->l->dtype);
          cp->code.push_back(di);
          break;
        case Node::BIN:
          go(nd->l.get());
          di.op = vop_of(nd->op);
          di.dtype = uint8_t(nd->l->dtype);
          di.mtype = mtype_of(nd->l->dtype);
          if (nd->r->kind == Node::LIT) {
            di.mode = RHS_IMM;
            di.imm = nd->r->imm;
          } else if (nd->r->kind == Node::COL) {
            di.mode = RHS_COL;
            di.slot = int16_t(pb->slot_of_column(nd->r->col));
          } else {
            go(nd->r.get());
            di.mode = RHS_STACK;
            // The evaluator keeps the TOP of the stack (here: the right operand) in its accumulator and
            // pops the operand below it as the second input, so a stack-mode instruction is emitted with
            // its operands exchanged: reverse subtract / divide, mirrored comparisons.
            switch (di.op) {
              case V_SUB: di.op = V_RSUB; break;
              case V_DIV: di.op = V_RDIV; break;
              case V_LT: di.op = V_GT; break;
              case V_LE: di.op = V_GE; break;
              case V_GT: di.op = V_LT; break;
              case V_GE: di.op = V_LE; break;
              default: break;
            }
            bump(-1);
          }
          cp->code.push_back(di);
          break;
      }
    }
  } em{this, &cp, &depth};
  em.go(st[0].get());
  cp.out_dtype = st[0]->dtype;
  if (st[0]->kind == Node::COL) {
    cp.is_plain_column = true;
    cp.plain_slot = cp.code[0].slot;
  }
  progs_.push_back(std::move(cp));
  return int(progs_.size()) - 1;
}

int ProgramBuilder::add_rowid() {
  CompiledProgram cp;
  DevInsn di;
  memset(&di, 0, sizeof(di));
  di.op = V_PUSH_ROWID;
  di.dtype = DFGPU_UINT64;
  di.mtype = MT_U;
  cp.code.push_back(di);
  cp.out_dtype = DFGPU_UINT64;
  cp.max_depth = 1;
  progs_.push_back(std::move(cp));
  return int(progs_.size()) - 1;
}

int ProgramBuilder::add_rowid_plus(unsigned long long bias) {
  CompiledProgram cp;
  DevInsn di;
  memset(&di, 0, sizeof(di));
  di.op = V_PUSH_ROWID;
  di.dtype = DFGPU_UINT64;
  di.mtype = MT_U;
  cp.code.push_back(di);
  memset(&di, 0, sizeof(di));
  di.op = V_ADD;
  di.mode = RHS_IMM;
  di.dtype = DFGPU_UINT64;
  di.mtype = MT_U;
  di.imm = bias;
  cp.code.push_back(di);
  cp.out_dtype = DFGPU_UINT64;
  cp.max_depth = 1;
  progs_.push_back(std::move(cp));
  return int(progs_.size()) - 1;
}

int ProgramBuilder::add_synthetic_column(const void* dptr, int dtype) {
  if (int(slots_.size()) >= kMaxCols) fail(DFGPU_ERR_NOT_IMPLEMENTED, "too many distinct columns");
  synth_.push_back(Synth{dptr, dtype});
  slots_.push_back(-int(synth_.size()));  // -1 - k
  CompiledProgram cp;
  DevInsn di;
  memset(&di, 0, sizeof(di));
  di.op = V_PUSH_COL;
  di.slot = int16_t(slots_.size() - 1);
  di.dtype = uint8_t(dtype);
  di.mtype = mtype_of(dtype);
  cp.code.push_back(di);
  cp.out_dtype = dtype;
  cp.max_depth = 1;
  progs_.push_back(std::move(cp));
  return int(progs_.size()) - 1;
}

void ProgramBuilder::finish(ProgramSet* out) const {
  memset(out, 0, sizeof(*out));
  if (int(progs_.size()) > kMaxProgs)
    fail(DFGPU_ERR_NOT_IMPLEMENTED, "more than " + std::to_string(kMaxProgs) + " expressions in one operator");
  int pc = 0, maxd = 1;
  for (size_t i = 0; i < progs_.size(); i++) {
    out->start[i] = uint8_t(pc);
    if (pc + int(progs_[i].code.size()) > kMaxInsn)
      fail(DFGPU_ERR_NOT_IMPLEMENTED, "expression programs exceed " + std::to_string(kMaxInsn) + " instructions");
    for (const auto& di : progs_[i].code) out->insn[pc++] = di;
    out->out_dtype[i] = uint8_t(progs_[i].out_dtype);
    out->nullable[i] = progs_[i].nullable ? 1 : 0;
    if (progs_[i].max_depth > maxd) maxd = progs_[i].max_depth;
  }
  out->f64_only = 1;
  for (int i = 0; i < pc; i++) {
    const DevInsn& di = out->insn[i];
    if (di.op == V_CAST || !(di.mtype == MT_F64 || di.mtype == MT_BOOL)) out->f64_only = 0;
  }
  out->start[progs_.size()] = uint8_t(pc);
  out->nprog = int(progs_.size());
  out->ncols = int(slots_.size());
  out->max_depth = maxd;
  for (size_t s = 0; s < slots_.size(); s++) {
    if (slots_[s] < 0) {
      const Synth& sy = synth_[size_t(-1 - slots_[s])];
      out->cols[s].ptr = sy.ptr;
      out->cols[s].validity = nullptr;
      out->cols[s].dtype = sy.dtype;
      continue;
    }
    const DevColumn& c = batch_->cols[size_t(slots_[s])];
    out->cols[s].ptr = c.values;
    out->cols[s].validity = c.null_count > 0 ? c.validity : nullptr;
    if (c.null_count > 0) out->has_nulls = 1;
    out->cols[s].dtype = c.dtype;
  }
}

}  // namespace dfgpu
