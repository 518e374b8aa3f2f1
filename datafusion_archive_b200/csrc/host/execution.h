// execution.h — C++ mirror of the reference's operator API on the hot path:
//   trait Relation            src/execution/relation.rs:27-32
//   trait DataSource          src/execution/datasource.rs:27-30   (+ CsvDataSource :33-58)
//   DataSourceRelation        src/execution/relation.rs:34-54
//   FilterRelation/ProjectRelation/AggregateRelation -> GPU relations calling the C ABI
//   ExecutionContext          src/execution/context.rs:33-197
#pragma once
#include <fstream>
#include <map>
#include <optional>

#include "sqlplanner.h"

namespace dfhost {

// ---- host-side Arrow arrays (what Relation::next hands to its consumer) -----------------------------
struct Array {
  DataType data_type = 0;
  int64_t len = 0;
  int64_t offset = 0;      // element offset into the buffers (ArrayData.offset)
  int64_t null_count = 0;
  // buffers: either owned (vectors) or borrowed (raw pointers into caller memory)
  std::vector<uint8_t> own_values, own_validity;
  std::vector<int32_t> own_offsets;
  const void* values = nullptr;
  const uint8_t* validity = nullptr;
  const int32_t* offsets = nullptr;
  int64_t values_bytes = 0;  // Utf8 byte buffer size
  std::shared_ptr<void> owner;  // keeps library-owned (pinned) result memory alive for zero-copy columns
  dfgpu_col view() const;    // borrowed Arrow view for the C ABI
};
using ArrayRef = std::shared_ptr<Array>;

struct RecordBatch {
  SchemaRef schema;
  std::vector<ArrayRef> columns;
  int64_t num_rows = 0;
};

struct Relation {
  virtual ~Relation() {}
  virtual std::optional<RecordBatch> next() = 0;
  virtual const SchemaRef& schema() const = 0;
};
using RelationRef = std::shared_ptr<Relation>;

struct DataSource {
  virtual ~DataSource() {}
  virtual const SchemaRef& schema() const = 0;
  virtual std::optional<RecordBatch> next() = 0;
};
using DataSourceRef = std::shared_ptr<DataSource>;

// CsvDataSource::new(filename, schema, batch_size): has_headers is hard-wired to true, exactly as
// the reference does (datasource.rs:41) — the first line is always dropped.
class CsvDataSource : public DataSource {
 public:
  CsvDataSource(const std::string& filename, SchemaRef schema, size_t batch_size);
  const SchemaRef& schema() const override { return schema_; }
  std::optional<RecordBatch> next() override;
 private:
  SchemaRef schema_;
  std::ifstream file_;
  size_t batch_size_;
  bool header_skipped_ = false;
  size_t line_no_ = 0;
};

// In-memory source over borrowed Arrow buffers, yielding batch_size-row slices (zero copy).
class MemoryDataSource : public DataSource {
 public:
  MemoryDataSource(SchemaRef schema, std::vector<ArrayRef> cols, size_t batch_size);
  const SchemaRef& schema() const override { return schema_; }
  std::optional<RecordBatch> next() override;
 private:
  SchemaRef schema_;
  std::vector<ArrayRef> cols_;
  int64_t nrows_ = 0, pos_ = 0, batch_size_ = 0;
};

class DataSourceRelation : public Relation {
 public:
  explicit DataSourceRelation(DataSourceRef ds) : schema_(ds->schema()), ds_(std::move(ds)) {}
  std::optional<RecordBatch> next() override { return ds_->next(); }
  const SchemaRef& schema() const override { return schema_; }
 private:
  SchemaRef schema_;
  DataSourceRef ds_;
};

// FilterRelation (+ ProjectRelation fused): src/execution/filter.rs:29-110, projection.rs:29-74
class GpuFilterProjectRelation : public Relation {
 public:
  // predicate may be null (projection only); proj empty = all input columns (FilterRelation alone)
  GpuFilterProjectRelation(dfgpu_ctx* gpu, RelationRef input, ExprRef predicate, std::vector<ExprRef> proj, SchemaRef schema);
  std::optional<RecordBatch> next() override;
  const SchemaRef& schema() const override { return schema_; }
 private:
  RecordBatch process(const RecordBatch& batch, const std::vector<ExprRef>& proj, const SchemaRef& out_schema);
  dfgpu_ctx* gpu_;
  RelationRef input_;
  ExprRef predicate_;
  std::vector<ExprRef> proj_;
  SchemaRef schema_;
};

// Row-range shard of a relation for one-process-per-GPU execution: rank g of G passes on rows
// [g*ceil(n/G), min(n, (g+1)*ceil(n/G))) of every batch (zero copy: Array offset / len).  Inserted above the
// TableScan by ExecutionContext::execute when a partition is set.
class ShardRelation : public Relation {
 public:
  ShardRelation(RelationRef input, int rank, int world) : input_(std::move(input)), rank_(rank), world_(world) {}
  std::optional<RecordBatch> next() override;
  const SchemaRef& schema() const override { return input_->schema(); }
 private:
  RelationRef input_;
  int rank_, world_;
};

// AggregateRelation: src/execution/aggregate.rs:38-61, 615-631.  `predicate` (may be null) is the expression
// of a Selection directly under the Aggregate (context.rs:126-139 builds FilterRelation there): it is fused
// into the scan kernel instead of materialising the filtered batch.
class GpuAggregateRelation : public Relation {
 public:
  GpuAggregateRelation(dfgpu_ctx* gpu, SchemaRef schema, RelationRef input, std::vector<ExprRef> group_expr, std::vector<ExprRef> aggr_expr,
                       ExprRef predicate = nullptr);
  std::optional<RecordBatch> next() override;
  const SchemaRef& schema() const override { return schema_; }
 private:
  dfgpu_ctx* gpu_;
  SchemaRef schema_;
  RelationRef input_;
  std::vector<ExprRef> group_expr_, aggr_expr_;
  ExprRef predicate_;
  bool end_of_results_ = false;
};

class ExecutionContext {
 public:
  explicit ExecutionContext(int device);  // ExecutionContext::new() + dfgpu_init
  ~ExecutionContext();
  RelationRef sql(const std::string& sql);                                    // context.rs:44-98
  void register_datasource(const std::string& name, DataSourceRef ds);        // context.rs:100-102
  RelationRef execute(const PlanRef& plan);                                   // context.rs:104-196
  PlanRef plan(const std::string& sql);                                       // parse + plan only
  dfgpu_ctx* gpu() const { return gpu_; }
  // One process (ExecutionContext) per GPU: attach this context to an NCCL communicator and make it work on
  // its row range of every table.  Aggregates then return the GLOBAL result on every rank (partial-aggregate
  // merge inside dfgpu_aggregate_finish); filter / project relations return this rank's rows, and the
  // rank-ordered concatenation of all ranks' outputs is the global output.  `nccl_unique_id`: 128 bytes from
  // dfgpu_comm_unique_id on rank 0.
  void set_partition(int rank, int world, const uint8_t* nccl_unique_id);
  int rank() const { return rank_; }
  int world() const { return world_; }
  bool verbose = false;  // the reference prints "Logical plan: ..." on every execute (context.rs:105)
 private:
  std::shared_ptr<std::map<std::string, DataSourceRef>> datasources_;
  dfgpu_ctx* gpu_ = nullptr;
  int rank_ = 0, world_ = 1;
};

// Expr name as RuntimeExpr::get_name reports it (expression.rs:230,312,322,407)
std::string runtime_expr_name(const Expr& e, const Schema& input_schema);

}  // namespace dfhost
