// sqlparser.h — SQL subset front-end.  The reference delegates parsing to crate sqlparser 0.2.1
// (Cargo.toml:34, wrapped by src/dfparser.rs:74); that crate is not under /root/reference, so this
// is a restatement of the grammar subset the planner consumes (src/sqlplanner.rs:46-375): one
// SELECT [list] [FROM ident] [WHERE e] [GROUP BY e,..] [HAVING e] [ORDER BY e [ASC|DESC],..] [LIMIT n].
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace dfhost {

enum class SQLOperator { Plus, Minus, Multiply, Divide, Modulus, Gt, Lt, GtEq, LtEq, Eq, NotEq, And, Or, Not, Like, NotLike };
enum class SQLType { Boolean, SmallInt, Int, BigInt, Float, Real, Double, Char, Varchar, Other };

struct ASTNode;
using ASTRef = std::shared_ptr<ASTNode>;
struct OrderByExpr { ASTRef expr; bool asc = true; };

struct ASTNode {
  enum Kind { SQLIdentifier, SQLWildcard, SQLLong, SQLDouble, SQLString, SQLBinaryExpr, SQLCast, SQLIsNull, SQLIsNotNull, SQLFunction, SQLSelect } kind = SQLIdentifier;
  std::string id;       // identifier / function name / string literal / unknown type name
  long long lval = 0;   // SQLLong
  double dval = 0;      // SQLDouble
  ASTRef left, right;   // binary; `left` = operand of cast / is-null
  SQLOperator op = SQLOperator::Eq;
  SQLType sql_type = SQLType::Other;
  std::vector<ASTRef> args;
  // SQLSelect
  std::vector<ASTRef> projection;
  ASTRef relation, selection, having, limit;
  bool has_group_by = false, has_order_by = false;
  std::vector<ASTRef> group_by;
  std::vector<OrderByExpr> order_by;
  std::string debug() const;
};

// Throws ExecutionError{DFGPU_ERR_GENERAL (ParserError), ...} on malformed input.
ASTRef parse_sql(const std::string& sql);

}  // namespace dfhost
