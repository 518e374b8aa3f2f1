// sqlplanner.cpp — see sqlplanner.h.  Rule-for-rule mirror of SqlToRel::sql_to_rel / sql_to_rex:
// literal typing (Long -> Int64, Double -> Float64), get_supertype + cast_to on both operands of
// every binary expression, aggregate detection by scanning the SELECT list, COUNT(1)/COUNT(*) ->
// COUNT(#0), output = group exprs ++ aggregate exprs.
#include "sqlplanner.h"

#include <algorithm>
#include <cctype>

namespace dfhost {

static std::string lower(std::string s) {
  for (auto& c : s) c = char(tolower((unsigned char)c));
  return s;
}

DataType convert_data_type(const ASTNode& n) {
  switch (n.sql_type) {
    case SQLType::Boolean: return DFGPU_BOOL;
    case SQLType::SmallInt: return DFGPU_INT16;
    case SQLType::Int: return DFGPU_INT32;
    case SQLType::BigInt: return DFGPU_INT64;
    case SQLType::Float: case SQLType::Real: case SQLType::Double: return DFGPU_FLOAT64;
    case SQLType::Char: case SQLType::Varchar: return DFGPU_UTF8;
    default: fail(DFGPU_ERR_NOT_IMPLEMENTED, "Unsupported SQL type " + n.id);
  }
}

Field expr_to_field(const Expr& e, const Schema& input_schema) {
  switch (e.kind) {
    case Expr::Column:
      if (e.index >= input_schema.fields.size()) fail(DFGPU_ERR_INVALID_COLUMN, "column index out of range");
      return input_schema.fields[e.index];
    case Expr::Literal: return Field{"lit", e.value.get_datatype(), true};
    case Expr::ScalarFunction: case Expr::AggregateFunction: return Field{e.name, e.data_type, true};
    case Expr::Cast: return Field{"cast", e.data_type, true};
    case Expr::BinaryExpr: {
      DataType st;
      if (!get_supertype(e.left->get_type(input_schema), e.right->get_type(input_schema), &st))
        fail(DFGPU_ERR_INTERNAL, "no common supertype (reference: unwrap() panic at sqlplanner.rs:422)");
      return Field{"binary_expr", st, true};
    }
    default: fail(DFGPU_ERR_NOT_IMPLEMENTED, "Cannot determine schema type for expression " + e.debug());
  }
}

std::vector<Field> exprlist_to_fields(const std::vector<ExprRef>& expr, const Schema& input_schema) {
  std::vector<Field> out;
  for (auto& e : expr) out.push_back(expr_to_field(*e, input_schema));
  return out;
}

PlanRef SqlToRel::sql_to_rel(const ASTRef& sql) const {
  switch (sql->kind) {
    case ASTNode::SQLSelect: {
      // parse the input relation so we have access to the row type
      PlanRef input;
      if (sql->relation) input = sql_to_rel(sql->relation);
      else {
        auto e = std::make_shared<LogicalPlan>();
        e->kind = LogicalPlan::EmptyRelation;
        e->schema_ = std::make_shared<Schema>();
        input = e;
      }
      const SchemaRef& input_schema = input->schema();

      // selection first
      PlanRef selection_plan;
      if (sql->selection) {
        auto s = std::make_shared<LogicalPlan>();
        s->kind = LogicalPlan::Selection;
        s->expr.push_back(sql_to_rex(sql->selection, *input_schema));
        s->input = input;
        selection_plan = s;
      }

      std::vector<ExprRef> expr;
      for (auto& e : sql->projection) expr.push_back(sql_to_rex(e, *input_schema));

      // collect aggregate expressions
      std::vector<ExprRef> aggr_expr;
      for (auto& e : expr)
        if (e->kind == Expr::AggregateFunction) aggr_expr.push_back(e);

      if (!aggr_expr.empty()) {
        PlanRef aggregate_input = selection_plan ? selection_plan : input;
        std::vector<ExprRef> group_expr;
        if (sql->has_group_by)
          for (auto& e : sql->group_by) group_expr.push_back(sql_to_rex(e, *input_schema));
        std::vector<ExprRef> all_fields = group_expr;
        for (auto& x : aggr_expr) all_fields.push_back(x);
        auto aggr_schema = std::make_shared<Schema>();
        aggr_schema->fields = exprlist_to_fields(all_fields, *input_schema);
        auto a = std::make_shared<LogicalPlan>();
        a->kind = LogicalPlan::Aggregate;
        a->input = aggregate_input;
        a->group_expr = group_expr;
        a->aggr_expr = aggr_expr;
        a->schema_ = aggr_schema;
        return a;
      }

      PlanRef projection_input = selection_plan ? selection_plan : input;
      auto projection_schema = std::make_shared<Schema>();
      projection_schema->fields = exprlist_to_fields(expr, *input_schema);
      auto proj = std::make_shared<LogicalPlan>();
      proj->kind = LogicalPlan::Projection;
      proj->expr = expr;
      proj->input = projection_input;
      proj->schema_ = projection_schema;

      if (sql->having) fail(DFGPU_ERR_GENERAL, "HAVING is not implemented yet");

      PlanRef order_by_plan = proj;
      if (sql->has_order_by) {
        auto s = std::make_shared<LogicalPlan>();
        s->kind = LogicalPlan::Sort;
        for (auto& o : sql->order_by) s->expr.push_back(Expr::sort(sql_to_rex(o.expr, *proj->schema()), o.asc));
        s->input = proj;
        s->schema_ = proj->schema();
        order_by_plan = s;
      }
      if (sql->limit) {
        if (sql->limit->kind != ASTNode::SQLLong) fail(DFGPU_ERR_GENERAL, "LIMIT parameter is not a number");
        auto l = std::make_shared<LogicalPlan>();
        l->kind = LogicalPlan::Limit;
        l->limit = size_t(sql->limit->lval);
        l->schema_ = order_by_plan->schema();
        l->input = order_by_plan;
        return l;
      }
      return order_by_plan;
    }
    case ASTNode::SQLIdentifier: {
      SchemaRef schema = schema_provider_->get_table_meta(sql->id);
      if (!schema) fail(DFGPU_ERR_GENERAL, "no schema found for table " + sql->id);
      auto t = std::make_shared<LogicalPlan>();
      t->kind = LogicalPlan::TableScan;
      t->schema_name = "default";
      t->table_name = sql->id;
      t->schema_ = schema;
      return t;
    }
    default: fail(DFGPU_ERR_EXECUTION, "sql_to_rel does not support this relation: " + sql->debug());
  }
}

ExprRef SqlToRel::sql_to_rex(const ASTRef& sql, const Schema& schema) const {
  switch (sql->kind) {
    case ASTNode::SQLLong: return Expr::literal(ScalarValue::Int64(sql->lval));
    case ASTNode::SQLDouble: return Expr::literal(ScalarValue::Float64(sql->dval));
    case ASTNode::SQLString: return Expr::literal(ScalarValue::Utf8(sql->id));
    case ASTNode::SQLIdentifier: {
      for (size_t i = 0; i < schema.fields.size(); i++)
        if (schema.fields[i].name == sql->id) return Expr::column(i);
      fail(DFGPU_ERR_EXECUTION, "Invalid identifier '" + sql->id + "' for schema " + schema.to_string());
    }
    case ASTNode::SQLWildcard:
      fail(DFGPU_ERR_NOT_IMPLEMENTED, "SQL wildcard operator is not supported in projection - please use explicit column names");
    case ASTNode::SQLCast: return Expr::cast(sql_to_rex(sql->left, schema), convert_data_type(*sql));
    case ASTNode::SQLIsNull: return Expr::is_null(sql_to_rex(sql->left, schema), false);
    case ASTNode::SQLIsNotNull: return Expr::is_null(sql_to_rex(sql->left, schema), true);
    case ASTNode::SQLBinaryExpr: {
      Operator op;
      switch (sql->op) {
        case SQLOperator::Gt: op = Operator::Gt; break;
        case SQLOperator::GtEq: op = Operator::GtEq; break;
        case SQLOperator::Lt: op = Operator::Lt; break;
        case SQLOperator::LtEq: op = Operator::LtEq; break;
        case SQLOperator::Eq: op = Operator::Eq; break;
        case SQLOperator::NotEq: op = Operator::NotEq; break;
        case SQLOperator::Plus: op = Operator::Plus; break;
        case SQLOperator::Minus: op = Operator::Minus; break;
        case SQLOperator::Multiply: op = Operator::Multiply; break;
        case SQLOperator::Divide: op = Operator::Divide; break;
        case SQLOperator::Modulus: op = Operator::Modulus; break;
        case SQLOperator::And: op = Operator::And; break;
        case SQLOperator::Or: op = Operator::Or; break;
        case SQLOperator::Not: op = Operator::Not; break;
        case SQLOperator::Like: op = Operator::Like; break;
        default: op = Operator::NotLike; break;
      }
      ExprRef left_expr = sql_to_rex(sql->left, schema), right_expr = sql_to_rex(sql->right, schema);
      DataType lt = left_expr->get_type(schema), rt = right_expr->get_type(schema), st;
      if (!get_supertype(lt, rt, &st))
        fail(DFGPU_ERR_GENERAL, std::string("No common supertype found for binary operator ") + operator_debug(op) + " with input types " +
                                    datatype_debug(lt) + " and " + datatype_debug(rt));
      return Expr::binary(left_expr->cast_to(st, schema), op, right_expr->cast_to(st, schema));
    }
    case ASTNode::SQLFunction: {
      const std::string lid = lower(sql->id);
      if (lid == "min" || lid == "max" || lid == "sum" || lid == "avg") {
        std::vector<ExprRef> rex_args;
        for (auto& a : sql->args) rex_args.push_back(sql_to_rex(a, schema));
        if (rex_args.empty()) fail(DFGPU_ERR_INTERNAL, "aggregate function without arguments (reference: index panic at sqlplanner.rs:320)");
        // return type is same as the argument type for these aggregate functions
        return Expr::aggregate(sql->id, rex_args, rex_args[0]->get_type(schema));
      }
      if (lid == "count") {
        std::vector<ExprRef> rex_args;
        for (auto& a : sql->args) {
          // COUNT(1) / COUNT(*) -> COUNT(first_column)
          if ((a->kind == ASTNode::SQLLong && a->lval == 1) || a->kind == ASTNode::SQLWildcard) rex_args.push_back(Expr::column(0));
          else rex_args.push_back(sql_to_rex(a, schema));
        }
        return Expr::aggregate(sql->id, rex_args, DFGPU_UINT64);
      }
      auto fm = schema_provider_->get_function_meta(sql->id);
      if (!fm) fail(DFGPU_ERR_GENERAL, "Invalid function '" + sql->id + "'");
      std::vector<ExprRef> safe_args;
      for (size_t i = 0; i < sql->args.size(); i++) {
        ExprRef a = sql_to_rex(sql->args[i], schema);
        if (i >= fm->args.size()) fail(DFGPU_ERR_INTERNAL, "too many function arguments (reference: index panic at sqlplanner.rs:356)");
        safe_args.push_back(a->cast_to(fm->args[i].data_type, schema));
      }
      return Expr::scalar_fn(sql->id, safe_args, fm->return_type);
    }
    default: fail(DFGPU_ERR_GENERAL, "Unsupported ast node " + sql->debug() + " in sqltorel");
  }
}

}  // namespace dfhost
