// sqlplanner.h — SQL AST -> LogicalPlan, mirroring SqlToRel (src/sqlplanner.rs:28-375).
#pragma once
#include <functional>

#include "logicalplan.h"
#include "sqlparser.h"

namespace dfhost {

struct FunctionMeta {  // src/logicalplan.rs:30-64
  std::string name;
  std::vector<Field> args;
  DataType return_type = 0;
};

struct SchemaProvider {  // trait SchemaProvider, src/sqlplanner.rs:28-31
  virtual ~SchemaProvider() {}
  virtual SchemaRef get_table_meta(const std::string& name) const = 0;
  virtual std::shared_ptr<FunctionMeta> get_function_meta(const std::string& name) const = 0;
};

class SqlToRel {
 public:
  explicit SqlToRel(std::shared_ptr<SchemaProvider> sp) : schema_provider_(std::move(sp)) {}
  PlanRef sql_to_rel(const ASTRef& sql) const;                     // sqlplanner.rs:46-209
  ExprRef sql_to_rex(const ASTRef& sql, const Schema& schema) const;  // sqlplanner.rs:212-375
 private:
  std::shared_ptr<SchemaProvider> schema_provider_;
};

DataType convert_data_type(const ASTNode& cast_node);              // sqlplanner.rs:379-394
Field expr_to_field(const Expr& e, const Schema& input_schema);    // sqlplanner.rs:396-428
std::vector<Field> exprlist_to_fields(const std::vector<ExprRef>& expr, const Schema& input_schema);

}  // namespace dfhost
