// memgraph_b200/csrc/pagerank_module.cpp -- the drop-in `pagerank` query module (host side, C++).
//
// Replaces mage/cpp/pagerank_module/pagerank_module.cpp of the reference behind the same ABI:
//   CALL pagerank.get(max_iterations = 100, damping_factor = 0.85, stop_epsilon = 1e-5, num_of_threads = 1)
//   YIELD node, rank                                                      (reference :116-155)
// Same procedure name, argument names / types / defaults / order, result fields, one row per
// vertex in vertex-iteration order, same error texts.  What differs is where the arithmetic runs:
// the graph is pulled ONCE through the mgp_graph iterators (reference CreatePageRankGraph :18-54),
// shipped to the GPU, iterated there (include/mgb200_pagerank.h), and the ranks come back as one
// double[N] that is turned into records (reference InsertPagerankRecord :69-84).
//
// Build: memgraph_b200/build.py links this file with the CUDA objects and a static CUDA runtime into
// memgraph_b200/_build/pagerank.so, whose only undefined symbols are mgp_* (plus libc/libstdc++),
// as dlopen(RTLD_NOW | RTLD_LOCAL) in src/query/procedure/module.cpp:861 requires.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>

#include "mgb200_pagerank.h"
#include "mgp_abi.h"

#ifdef MGB200_USE_REFERENCE_MGP_HEADER
// contract-check build against the reference header: keep the optional imports weak there too
extern "C" {
enum mgp_error mgp_result_reserve(struct mgp_result *res, size_t n) __attribute__((weak));
enum mgp_error mgp_log(enum mgp_log_level log_level, const char *output) __attribute__((weak));
int mgp_must_abort(struct mgp_graph *graph) __attribute__((weak));
}
#endif

namespace {

constexpr const char *kProcedureGet = "get";
constexpr const char *kFieldNode = "node";
constexpr const char *kFieldRank = "rank";
constexpr const char *kArgMaxIterations = "max_iterations";
constexpr const char *kArgDampingFactor = "damping_factor";
constexpr const char *kArgStopEpsilon = "stop_epsilon";
constexpr const char *kArgNumThreads = "num_of_threads";

// ---- mgp_error -> exception, with the texts a user of the reference module would see ---------------
// (include/_mgp.hpp:25-56 maps codes to mg_exception types; include/mg_exceptions.hpp holds the texts.)
struct ModuleError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

const char *ErrorText(mgp_error code) {
  switch (code) {
    case mgp_error::MGP_ERROR_UNKNOWN_ERROR: return "Unknown exception!";
    case mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE:
      return "Could not allocate memory. For more details please visit https://memgr.ph/memory-control";
    case mgp_error::MGP_ERROR_INSUFFICIENT_BUFFER: return "Buffer is not sufficient to process procedure!";
    case mgp_error::MGP_ERROR_OUT_OF_RANGE: return "Index out of range!";
    case mgp_error::MGP_ERROR_LOGIC_ERROR: return "Logic exception, check the procedure signature!";
    case mgp_error::MGP_ERROR_DELETED_OBJECT: return "Object is deleted!";
    case mgp_error::MGP_ERROR_INVALID_ARGUMENT: return "Invalid argument!";
    case mgp_error::MGP_ERROR_KEY_ALREADY_EXISTS: return "Key you are trying to set already exists!";
    case mgp_error::MGP_ERROR_IMMUTABLE_OBJECT: return "Object you are trying to change is immutable!";
    case mgp_error::MGP_ERROR_VALUE_CONVERSION: return "Error in value conversion!";
    case mgp_error::MGP_ERROR_SERIALIZATION_ERROR: return "Error in serialization!";
    case mgp_error::MGP_ERROR_NOT_YET_IMPLEMENTED: return "Not yet implemented!";
    default: return nullptr;  // NO_ERROR and AUTHORIZATION_ERROR fall through in the reference as well
  }
}

inline void Check(mgp_error code) {
  if (const char *text = ErrorText(code)) throw ModuleError(text);
}

// Owning wrappers so no iterator / value leaks on any exit path (ownership rules: mg_procedure.h:887-920).
struct VerticesIt {
  mgp_vertices_iterator *it = nullptr;
  VerticesIt(mgp_graph *graph, mgp_memory *memory) { Check(mgp_graph_iter_vertices(graph, memory, &it)); }
  ~VerticesIt() {
    if (it) mgp_vertices_iterator_destroy(it);
  }
  mgp_vertex *Get() {
    mgp_vertex *v = nullptr;
    Check(mgp_vertices_iterator_get(it, &v));
    return v;
  }
  mgp_vertex *Next() {
    mgp_vertex *v = nullptr;
    Check(mgp_vertices_iterator_next(it, &v));
    return v;
  }
  VerticesIt(const VerticesIt &) = delete;
  VerticesIt &operator=(const VerticesIt &) = delete;
};

struct OutEdgesIt {
  mgp_edges_iterator *it = nullptr;
  OutEdgesIt(mgp_vertex *v, mgp_memory *memory) { Check(mgp_vertex_iter_out_edges(v, memory, &it)); }
  ~OutEdgesIt() {
    if (it) mgp_edges_iterator_destroy(it);
  }
  mgp_edge *Get() {
    mgp_edge *e = nullptr;
    Check(mgp_edges_iterator_get(it, &e));
    return e;
  }
  mgp_edge *Next() {
    mgp_edge *e = nullptr;
    Check(mgp_edges_iterator_next(it, &e));
    return e;
  }
  OutEdgesIt(const OutEdgesIt &) = delete;
  OutEdgesIt &operator=(const OutEdgesIt &) = delete;
};

struct OwnedValue {
  mgp_value *v = nullptr;
  ~OwnedValue() {
    if (v) mgp_value_destroy(v);
  }
};

// gid -> dense id, open addressing; a gid that was never visited resolves to dense id 0, which is
// what the reference's unordered_map::operator[] lookup yields (pagerank_module.cpp:50).
class GidMap {
 public:
  void Reserve(std::size_t n) {
    std::size_t cap = 16;
    while (cap < n * 2 + 2) cap <<= 1;
    keys_.assign(cap, 0);
    vals_.assign(cap, kEmpty);
    mask_ = cap - 1;
    size_ = 0;
  }
  void Put(int64_t gid, uint64_t dense) {
    if ((size_ + 1) * 2 > keys_.size()) Grow();
    Insert(gid, dense);
  }
  uint64_t GetOrZero(int64_t gid) const {
    if (keys_.empty()) return 0;
    std::size_t i = Hash(gid) & mask_;
    while (vals_[i] != kEmpty) {
      if (keys_[i] == gid) return vals_[i];
      i = (i + 1) & mask_;
    }
    return 0;
  }

 private:
  static constexpr uint64_t kEmpty = ~0ull;
  static std::size_t Hash(int64_t gid) {
    uint64_t z = static_cast<uint64_t>(gid) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return static_cast<std::size_t>(z ^ (z >> 31));
  }
  void Insert(int64_t gid, uint64_t dense) {
    std::size_t i = Hash(gid) & mask_;
    while (vals_[i] != kEmpty) {
      if (keys_[i] == gid) {
        vals_[i] = dense;  // same gid visited again: last assignment wins, as with operator[]
        return;
      }
      i = (i + 1) & mask_;
    }
    keys_[i] = gid;
    vals_[i] = dense;
    ++size_;
  }
  void Grow() {
    std::vector<int64_t> old_keys;
    std::vector<uint64_t> old_vals;
    old_keys.swap(keys_);
    old_vals.swap(vals_);
    const std::size_t cap = old_keys.empty() ? 16 : old_keys.size() * 2;
    keys_.assign(cap, 0);
    vals_.assign(cap, kEmpty);
    mask_ = cap - 1;
    size_ = 0;
    for (std::size_t i = 0; i < old_keys.size(); ++i)
      if (old_vals[i] != kEmpty) Insert(old_keys[i], old_vals[i]);
  }
  std::vector<int64_t> keys_;
  std::vector<uint64_t> vals_;
  std::size_t mask_ = 0, size_ = 0;
};

struct HostGraph {
  std::vector<int64_t> gid_of_dense;  // reference id_to_memgraph (:42)
  std::vector<uint64_t> from, to;     // dense (source, destination) per edge, in iteration order
};

bool MustAbort(mgp_graph *graph) { return mgp_must_abort != nullptr && mgp_must_abort(graph) != 0; }

// The reference's ingest (CreatePageRankGraph :18-54): one vertex iterator, one out-edge iterator per
// vertex, dense id = visit index, destinations remapped afterwards.
HostGraph PullGraph(mgp_graph *graph, mgp_memory *memory) {
  HostGraph hg;
  std::size_t approx_vertices = 0, approx_edges = 0;
  Check(mgp_graph_approximate_vertex_count(graph, &approx_vertices));
  Check(mgp_graph_approximate_edge_count(graph, &approx_edges));
  hg.gid_of_dense.reserve(approx_vertices);
  hg.from.reserve(approx_edges);
  std::vector<int64_t> to_gid;
  to_gid.reserve(approx_edges);
  GidMap dense_of_gid;
  dense_of_gid.Reserve(approx_vertices);

  VerticesIt vertices(graph, memory);
  for (mgp_vertex *source = vertices.Get(); source != nullptr; source = vertices.Next()) {
    mgp_vertex_id source_id{};
    Check(mgp_vertex_get_id(source, &source_id));
    const uint64_t source_dense = hg.gid_of_dense.size();
    {
      OutEdgesIt edges(source, memory);
      for (mgp_edge *edge = edges.Get(); edge != nullptr; edge = edges.Next()) {
        mgp_vertex *destination = nullptr;
        Check(mgp_edge_get_to(edge, &destination));
        mgp_vertex_id destination_id{};
        Check(mgp_vertex_get_id(destination, &destination_id));
        hg.from.push_back(source_dense);
        to_gid.push_back(destination_id.as_int);
      }
    }
    dense_of_gid.Put(source_id.as_int, source_dense);
    hg.gid_of_dense.push_back(source_id.as_int);
    if ((source_dense & 0xFFF) == 0 && MustAbort(graph)) throw ModuleError("pagerank.get aborted by the host");
  }
  hg.to.resize(to_gid.size());
  for (std::size_t e = 0; e < to_gid.size(); ++e) hg.to[e] = dense_of_gid.GetOrZero(to_gid[e]);
  // Reference quirk kept: `from` was recorded as the visit index, but the reference remaps the SOURCE
  // gid through the map as well (:50); the two differ only if a gid is visited twice, where the map
  // holds the later index.
  for (std::size_t e = 0; e < hg.from.size(); ++e)
    hg.from[e] = dense_of_gid.GetOrZero(hg.gid_of_dense[hg.from[e]]);
  return hg;
}

int AbortTrampoline(void *user) { return MustAbort(static_cast<mgp_graph *>(user)) ? 1 : 0; }

std::vector<double> ComputeRanks(const HostGraph &hg, mgp_graph *graph, int64_t max_iterations, double damping_factor,
                                 double stop_epsilon, int64_t num_threads) {
  // pagerank.hpp:109 takes uint32_t, pagerank.cpp:195 clamps to hardware_concurrency, :63-65 rejects 0.
  const uint32_t threads = static_cast<uint32_t>(num_threads);
  if (threads == 0) throw ModuleError(MGB200_MSG_ZERO_THREADS);
  const uint64_t n = hg.gid_of_dense.size();
  std::vector<double> ranks(n);
  if (n == 0) return ranks;
  // MGB200_GPUS=P (2..8): vertex-partition the call over GPUs 0..P-1 (device-side exchange over NVLink)
  const char *gpus_env = std::getenv("MGB200_GPUS");
  const int gpus = gpus_env ? std::atoi(gpus_env) : 1;
  if (gpus > 1) {
    uint64_t iterations = 0;
    const int rc = mgb200_parallel_iterative_pagerank_multi(
        n, hg.from.size(), hg.from.data(), hg.to.data(), static_cast<uint64_t>(max_iterations), damping_factor,
        stop_epsilon, threads, static_cast<uint32_t>(gpus), nullptr, ranks.data(), &iterations);
    if (rc != MGB200_OK) throw ModuleError(mgb200_last_error());
    return ranks;
  }
  const char *dev_env = std::getenv("MGB200_DEVICE");
  const int device = dev_env ? std::atoi(dev_env) : 0;
  mgb200_graph *dg = nullptr;
  int rc = mgb200_graph_create_host(device, n, hg.from.size(), hg.from.data(), hg.to.data(), 0, 1, &dg);
  if (rc != MGB200_OK) throw ModuleError(mgb200_last_error());
  mgb200_run_params params{};
  params.max_iterations = static_cast<uint64_t>(max_iterations);  // int64 -> size_t wrap, reference :95
  params.damping_factor = damping_factor;
  params.stop_epsilon = stop_epsilon;
  params.should_abort = AbortTrampoline;
  params.abort_user = graph;
  mgb200_run_stats stats{};
  rc = mgb200_pagerank_run(dg, &params, ranks.data(), &stats);
  const std::string message = rc != MGB200_OK ? mgb200_last_error() : "";
  mgb200_graph_destroy(dg);
  if (rc != MGB200_OK) throw ModuleError(message);
  return ranks;
}

// One (node, rank) row (reference InsertPagerankRecord :69-84).  Returns false when the vertex has
// vanished in analytical mode (row skipped), throws "Invalid ID!" in transactional mode.
void EmitRecord(mgp_graph *graph, mgp_result *result, mgp_memory *memory, int64_t gid, double rank) {
  mgp_vertex *vertex = nullptr;
  Check(mgp_graph_get_vertex_by_id(graph, mgp_vertex_id{gid}, memory, &vertex));
  if (vertex == nullptr) {
    int transactional = 0;
    Check(mgp_graph_is_transactional(graph, &transactional));
    if (transactional) throw ModuleError("Invalid ID!");
    return;
  }
  OwnedValue node;
  {
    const mgp_error e = mgp_value_make_vertex(vertex, &node.v);  // takes ownership of `vertex` on success
    if (e != mgp_error::MGP_ERROR_NO_ERROR) Check(e);
  }
  mgp_result_record *record = nullptr;
  Check(mgp_result_new_record(result, &record));
  if (record == nullptr)
    throw ModuleError("Not enough memory! For more details please visit https://memgr.ph/memory-control");
  Check(mgp_result_record_insert(record, kFieldNode, node.v));
  OwnedValue value;
  Check(mgp_value_make_double(rank, memory, &value.v));
  Check(mgp_result_record_insert(record, kFieldRank, value.v));
}

void PagerankGet(mgp_list *args, mgp_graph *graph, mgp_result *result, mgp_memory *memory) {
  try {
    mgp_value *arg = nullptr;
    int64_t max_iterations = 0, num_threads = 0;
    double damping_factor = 0.0, stop_epsilon = 0.0;
    Check(mgp_list_at(args, 0, &arg));
    Check(mgp_value_get_int(arg, &max_iterations));
    Check(mgp_list_at(args, 1, &arg));
    Check(mgp_value_get_double(arg, &damping_factor));
    Check(mgp_list_at(args, 2, &arg));
    Check(mgp_value_get_double(arg, &stop_epsilon));
    Check(mgp_list_at(args, 3, &arg));
    Check(mgp_value_get_int(arg, &num_threads));

    const HostGraph hg = PullGraph(graph, memory);
    const std::vector<double> ranks =
        ComputeRanks(hg, graph, max_iterations, damping_factor, stop_epsilon, num_threads);

    if (mgp_result_reserve != nullptr && !ranks.empty()) (void)mgp_result_reserve(result, ranks.size());
    for (std::size_t dense = 0; dense < ranks.size(); ++dense)
      EmitRecord(graph, result, memory, hg.gid_of_dense[dense], ranks[dense]);
  } catch (const std::exception &e) {
    // Nothing may escape the callback (reference :108-112); the engine turns the message into
    // QueryRuntimeException("pagerank.get: <msg>") (src/query/plan/operator.cpp:7929-7932).
    (void)mgp_result_set_error_msg(result, e.what());
  } catch (...) {
    (void)mgp_result_set_error_msg(result, "Unknown exception!");
  }
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int mgp_init_module(struct mgp_module *module, struct mgp_memory *memory) {
  OwnedValue max_iterations, damping_factor, stop_epsilon, num_threads;  // destroyed on every path (:139-152)
  try {
    mgp_proc *proc = nullptr;
    Check(mgp_module_add_read_procedure(module, kProcedureGet, PagerankGet, &proc));
    Check(mgp_value_make_int(100, memory, &max_iterations.v));
    Check(mgp_value_make_double(0.85, memory, &damping_factor.v));
    Check(mgp_value_make_double(1e-5, memory, &stop_epsilon.v));
    Check(mgp_value_make_int(1, memory, &num_threads.v));
    mgp_type *type_int = nullptr, *type_float = nullptr, *type_node = nullptr;
    Check(mgp_type_int(&type_int));
    Check(mgp_type_float(&type_float));
    Check(mgp_type_node(&type_node));
    Check(mgp_proc_add_opt_arg(proc, kArgMaxIterations, type_int, max_iterations.v));
    Check(mgp_proc_add_opt_arg(proc, kArgDampingFactor, type_float, damping_factor.v));
    Check(mgp_proc_add_opt_arg(proc, kArgStopEpsilon, type_float, stop_epsilon.v));
    Check(mgp_proc_add_opt_arg(proc, kArgNumThreads, type_int, num_threads.v));
    Check(mgp_proc_add_result(proc, kFieldNode, type_node));
    Check(mgp_proc_add_result(proc, kFieldRank, type_float));
  } catch (...) {
    return 1;
  }
  return 0;
}

__attribute__((visibility("default"))) int mgp_shutdown_module(void) { return 0; }

}  // extern "C"
