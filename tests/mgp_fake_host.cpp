// tests/mgp_fake_host.cpp -- a minimal in-memory HOST for Memgraph query modules.
//
// TEST INFRASTRUCTURE.  The real host (memgraph: src/query/procedure/mg_procedure_impl.cpp, module.cpp,
// src/query/plan/operator.cpp) cannot be built here (35 conan packages, no network), so this file plays
// its part for exactly the slice of the ABI a PageRank module touches:
//   * implements the 30 mgp_* functions the reference pagerank.so imports, plus mgp_must_abort,
//     mgp_result_reserve, mgp_log and mgp_vertex_destroy, with the behaviour documented in
//     SURVEY Appendix A (each function cites the host implementation it mimics);
//   * loads a module the way SharedLibraryModule::Load does (module.cpp:855-918): dlopen(RTLD_NOW |
//     RTLD_LOCAL), dlsym mgp_init_module (required, non-zero => load fails), dlsym mgp_shutdown_module
//     (optional); Close calls shutdown then dlclose (:920-939);
//   * calls a procedure the way CallCustomProcedure does (operator.cpp:7703-7800): argument count
//     check, STRICT type check (INTEGER != FLOAT), defaults for missing optional args, then
//     proc.cb(args, graph, result, memory) on the calling thread; an error message set by the module
//     becomes "<module>.<proc>: <msg>" (operator.cpp:7929-7932).
// It is loaded into the test process with RTLD_GLOBAL so that a module's undefined mgp_* symbols bind
// to it, like they bind to the memgraph executable's dynamic symbol table (include/mg_procedure.syms).
// The same host loads the REFERENCE's own pagerank module (oracle/_ref/pagerank_reference.so), which
// is how the host itself is validated (BASELINE config #1).
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cctype>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "mgp_abi.h"

// ---- host-side object model --------------------------------------------------------------------------

static std::atomic<long> g_live_objects{0};  // values, owned vertices and iterators handed to the module

struct mgp_memory {
  long allocations = 0;
};

struct FakeGraphData {
  std::vector<int64_t> gids;                 // ascending, like Vertices(view) over in-memory storage
  std::vector<std::vector<uint32_t>> out;    // out[v] = indices of destination vertices, insertion order
  std::vector<std::vector<double>> out_prop; // optional: one numeric edge property per out-edge (NaN = property absent)
  std::string prop_name;                     // its name (fh_graph_set_edge_property)
  bool prop_as_int = false;                  // hand the property out as an INTEGER value (modules must convert)
  std::unordered_map<int64_t, uint32_t> index_of;
  std::vector<char> hidden;                  // hidden[v] != 0: FindVertex fails at emission (vertex vanished mid-call, analytical mode)
  uint64_t edge_count = 0;
  int transactional = 1;
  std::atomic<int> abort_flag{0};
  std::atomic<long> abort_polls{0};
};

struct mgp_graph {
  FakeGraphData *data;
};

struct mgp_vertex {
  mgp_graph *graph;
  uint32_t index;
  bool owned;  // created by mgp_graph_get_vertex_by_id (counts as a live object)
};

struct mgp_edge {
  mgp_vertex from;
  mgp_vertex to;
  double prop = std::numeric_limits<double>::quiet_NaN();  // the edge's numeric property, NaN = absent
};

enum class ValueKind { Null, Int, Double, Vertex, String, List };

struct mgp_list;
struct mgp_value {
  ValueKind kind = ValueKind::Null;
  int64_t i = 0;
  double d = 0.0;
  mgp_vertex *vertex = nullptr;  // owned when kind == Vertex
  bool counted = false;          // allocated through the module-facing API
  std::string s;                         // kind == String
  std::shared_ptr<mgp_list> list;        // kind == List
};

struct mgp_list {
  std::vector<mgp_value> items;
};

struct mgp_type {
  std::string name;
  ValueKind kind;
  const mgp_type *element = nullptr;  // kind == List
};

struct mgp_vertices_iterator {
  mgp_graph *graph;
  uint32_t pos;
  mgp_vertex current;
};

struct mgp_edges_iterator {
  mgp_graph *graph;
  uint32_t source;
  std::vector<uint32_t> snapshot;  // COPY of the out-edge list at creation (mg_procedure_impl.cpp:2952-3040)
  size_t pos;
  mgp_edge current;
};

struct ProcArg {
  std::string name;
  mgp_type *type;
  mgp_value default_value;
  bool required = false;
};

struct mgp_proc {
  std::string name;
  mgp_proc_cb cb;
  std::vector<ProcArg> opt_args;
  std::vector<std::pair<std::string, mgp_type *>> results;
};

struct mgp_module {
  std::map<std::string, mgp_proc> procs;
};

struct mgp_result_record {
  struct mgp_result *owner;
  std::vector<mgp_value> fields;  // Null-prefilled, one per signature field
};

struct mgp_result {
  mgp_proc *proc;
  std::vector<std::unique_ptr<mgp_result_record>> rows;
  bool has_error = false;
  std::string error;
};

namespace {

mgp_type g_type_int{"INTEGER", ValueKind::Int};
mgp_type g_type_float{"FLOAT", ValueKind::Double};
mgp_type g_type_node{"NODE", ValueKind::Vertex};
mgp_type g_type_string{"STRING", ValueKind::String};
mgp_type g_type_list_node{"LIST OF NODE", ValueKind::List, &g_type_node};
mgp_type g_type_list_float{"LIST OF FLOAT", ValueKind::List, &g_type_float};
mgp_type g_type_list_int{"LIST OF INTEGER", ValueKind::List, &g_type_int};

bool ValidIdentifier(const char *s) {  // [_[:alpha:]][_[:alnum:]]*  (mg_procedure_impl.cpp:4832-4843)
  if (!s || !*s) return false;
  if (!(std::isalpha(static_cast<unsigned char>(*s)) || *s == '_')) return false;
  for (const char *p = s + 1; *p; ++p)
    if (!(std::isalnum(static_cast<unsigned char>(*p)) || *p == '_')) return false;
  return true;
}

constexpr mgp_error OK = mgp_error::MGP_ERROR_NO_ERROR;

}  // namespace

// ---- the mgp_* ABI ----------------------------------------------------------------------------------------

extern "C" {

// values (mg_procedure_impl.cpp:1211-1243, 1294-1302)
void mgp_value_destroy(mgp_value *val) {
  if (!val) return;
  if (val->kind == ValueKind::Vertex && val->vertex) {
    if (val->vertex->owned) g_live_objects--;
    delete val->vertex;
  }
  if (val->counted) g_live_objects--;
  delete val;
}
mgp_error mgp_value_make_int(int64_t v, mgp_memory *memory, mgp_value **result) {
  if (memory) memory->allocations++;
  auto *x = new (std::nothrow) mgp_value();
  if (!x) return mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE;
  x->kind = ValueKind::Int;
  x->i = v;
  x->counted = true;
  g_live_objects++;
  *result = x;
  return OK;
}
mgp_error mgp_value_make_double(double v, mgp_memory *memory, mgp_value **result) {
  if (memory) memory->allocations++;
  auto *x = new (std::nothrow) mgp_value();
  if (!x) return mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE;
  x->kind = ValueKind::Double;
  x->d = v;
  x->counted = true;
  g_live_objects++;
  *result = x;
  return OK;
}
mgp_error mgp_value_make_vertex(mgp_vertex *v, mgp_value **result) {  // takes ownership of v
  auto *x = new (std::nothrow) mgp_value();
  if (!x) return mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE;
  x->kind = ValueKind::Vertex;
  x->vertex = v;
  x->counted = true;
  g_live_objects++;
  *result = x;
  return OK;
}
mgp_error mgp_value_get_int(mgp_value *val, int64_t *result) {  // unchecked union read in the real host
  *result = val->i;
  return OK;
}
mgp_error mgp_value_get_double(mgp_value *val, double *result) {
  *result = val->d;
  return OK;
}
mgp_error mgp_list_at(mgp_list *list, size_t index, mgp_value **result) {  // :1385-1394
  if (index >= list->items.size()) return mgp_error::MGP_ERROR_OUT_OF_RANGE;
  *result = &list->items[index];
  return OK;
}

mgp_error mgp_list_size(mgp_list *list, size_t *result) {  // mg_procedure.h:554
  *result = list->items.size();
  return OK;
}
mgp_error mgp_value_get_list(mgp_value *val, mgp_list **result) {  // :451, unchecked in the real host
  *result = val->list.get();
  return OK;
}
mgp_error mgp_value_get_string(mgp_value *val, const char **result) {  // :446
  *result = val->s.c_str();
  return OK;
}
mgp_error mgp_value_make_string(const char *v, mgp_memory *memory, mgp_value **result) {  // :248 copies the string
  if (memory) memory->allocations++;
  auto *x = new (std::nothrow) mgp_value();
  if (!x) return mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE;
  x->kind = ValueKind::String;
  x->s = v ? v : "";
  x->counted = true;
  g_live_objects++;
  *result = x;
  return OK;
}
mgp_error mgp_value_get_type(mgp_value *val, mgp_value_type *result) {  // :353
  switch (val->kind) {
    case ValueKind::Null: *result = mgp_value_type::MGP_VALUE_TYPE_NULL; break;
    case ValueKind::Int: *result = mgp_value_type::MGP_VALUE_TYPE_INT; break;
    case ValueKind::Double: *result = mgp_value_type::MGP_VALUE_TYPE_DOUBLE; break;
    case ValueKind::Vertex: *result = mgp_value_type::MGP_VALUE_TYPE_VERTEX; break;
    case ValueKind::String: *result = mgp_value_type::MGP_VALUE_TYPE_STRING; break;
    case ValueKind::List: *result = mgp_value_type::MGP_VALUE_TYPE_LIST; break;
  }
  return OK;
}
// :967 -- a NEW value the caller destroys: the property, or Null when the edge does not have it
mgp_error mgp_edge_get_property(mgp_edge *e, const char *property_name, mgp_memory *memory, mgp_value **result) {
  if (memory) memory->allocations++;
  auto *x = new (std::nothrow) mgp_value();
  if (!x) return mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE;
  const FakeGraphData *d = e->from.graph->data;
  if (property_name && d->prop_name == property_name && e->prop == e->prop) {
    if (d->prop_as_int) {
      x->kind = ValueKind::Int;
      x->i = static_cast<int64_t>(e->prop);
    } else {
      x->kind = ValueKind::Double;
      x->d = e->prop;
    }
  }
  x->counted = true;
  g_live_objects++;
  *result = x;
  return OK;
}

// result rows (:2165-2216)
mgp_error mgp_result_set_error_msg(mgp_result *res, const char *msg) {
  res->has_error = true;
  res->error = msg ? msg : "";
  return OK;
}
mgp_error mgp_result_new_record(mgp_result *res, mgp_result_record **result) {
  auto rec = std::make_unique<mgp_result_record>();
  rec->owner = res;
  rec->fields.resize(res->proc->results.size());
  *result = rec.get();
  res->rows.push_back(std::move(rec));
  return OK;
}
mgp_error mgp_result_reserve(mgp_result *res, size_t n) {
  res->rows.reserve(n);
  return OK;
}
mgp_error mgp_result_record_insert(mgp_result_record *record, const char *field_name, mgp_value *val) {
  auto &sig = record->owner->proc->results;
  for (size_t i = 0; i < sig.size(); ++i) {
    if (sig[i].first != field_name) continue;
    if (val->kind != sig[i].second->kind) return mgp_error::MGP_ERROR_LOGIC_ERROR;
    mgp_value copy;
    copy.kind = val->kind;
    copy.i = val->i;
    copy.d = val->d;
    if (val->kind == ValueKind::Vertex) {
      copy.i = val->vertex->graph->data->gids[val->vertex->index];  // the row keeps the gid
    }
    record->fields[i] = copy;
    return OK;
  }
  return mgp_error::MGP_ERROR_OUT_OF_RANGE;
}

// vertices / edges (:2278-2282, 2952-3040, 3115, 3290-3312, 3682-3685, 4660-4740)
mgp_error mgp_vertex_get_id(mgp_vertex *v, mgp_vertex_id *result) {
  result->as_int = v->graph->data->gids[v->index];
  return OK;
}
void mgp_vertex_destroy(mgp_vertex *v) {
  if (!v) return;
  if (v->owned) g_live_objects--;
  delete v;
}
mgp_error mgp_vertex_iter_out_edges(mgp_vertex *v, mgp_memory *memory, mgp_edges_iterator **result) {
  if (memory) memory->allocations++;
  auto *it = new (std::nothrow) mgp_edges_iterator();
  if (!it) return mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE;
  it->graph = v->graph;
  it->source = v->index;
  it->snapshot = v->graph->data->out[v->index];
  it->pos = 0;
  if (!it->snapshot.empty()) {
    it->current = mgp_edge{{it->graph, it->source, false}, {it->graph, it->snapshot[0], false}};
    if (!v->graph->data->out_prop.empty()) it->current.prop = v->graph->data->out_prop[it->source][0];
  }
  g_live_objects++;
  *result = it;
  return OK;
}
void mgp_edges_iterator_destroy(mgp_edges_iterator *it) {
  if (!it) return;
  g_live_objects--;
  delete it;
}
mgp_error mgp_edges_iterator_get(mgp_edges_iterator *it, mgp_edge **result) {
  *result = it->pos < it->snapshot.size() ? &it->current : nullptr;
  return OK;
}
mgp_error mgp_edges_iterator_next(mgp_edges_iterator *it, mgp_edge **result) {
  if (it->pos < it->snapshot.size()) it->pos++;
  if (it->pos < it->snapshot.size()) {
    it->current = mgp_edge{{it->graph, it->source, false}, {it->graph, it->snapshot[it->pos], false}};
    if (!it->graph->data->out_prop.empty()) it->current.prop = it->graph->data->out_prop[it->source][it->pos];
    *result = &it->current;
  } else {
    *result = nullptr;
  }
  return OK;
}
mgp_error mgp_edge_get_to(mgp_edge *e, mgp_vertex **result) {
  *result = &e->to;
  return OK;
}
mgp_error mgp_graph_get_vertex_by_id(mgp_graph *g, mgp_vertex_id id, mgp_memory *memory, mgp_vertex **result) {
  if (memory) memory->allocations++;
  auto found = g->data->index_of.find(id.as_int);
  if (found == g->data->index_of.end() || g->data->hidden[found->second]) {
    *result = nullptr;
    return OK;
  }
  auto *v = new (std::nothrow) mgp_vertex{g, found->second, true};
  if (!v) return mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE;
  g_live_objects++;
  *result = v;
  return OK;
}
mgp_error mgp_graph_is_transactional(mgp_graph *graph, int *result) {
  *result = graph->data->transactional;
  return OK;
}
mgp_error mgp_graph_iter_vertices(mgp_graph *g, mgp_memory *memory, mgp_vertices_iterator **result) {
  if (memory) memory->allocations++;
  auto *it = new (std::nothrow) mgp_vertices_iterator{g, 0, mgp_vertex{g, 0, false}};
  if (!it) return mgp_error::MGP_ERROR_UNABLE_TO_ALLOCATE;
  g_live_objects++;
  *result = it;
  return OK;
}
void mgp_vertices_iterator_destroy(mgp_vertices_iterator *it) {
  if (!it) return;
  g_live_objects--;
  delete it;
}
mgp_error mgp_vertices_iterator_get(mgp_vertices_iterator *it, mgp_vertex **result) {
  *result = it->pos < it->graph->data->gids.size() ? &it->current : nullptr;
  return OK;
}
mgp_error mgp_vertices_iterator_next(mgp_vertices_iterator *it, mgp_vertex **result) {
  const size_t n = it->graph->data->gids.size();
  if (it->pos < n) it->pos++;
  if (it->pos < n) {
    it->current = mgp_vertex{it->graph, it->pos, false};
    *result = &it->current;
  } else {
    *result = nullptr;
  }
  return OK;
}
mgp_error mgp_graph_approximate_vertex_count(mgp_graph *graph, size_t *result) {
  *result = graph->data->gids.size();
  return OK;
}
mgp_error mgp_graph_approximate_edge_count(mgp_graph *graph, size_t *result) {
  *result = graph->data->edge_count;
  return OK;
}

// types and registration (:4761-4781, 4832-4865, 4922-5015)
mgp_error mgp_type_int(mgp_type **result) {
  *result = &g_type_int;
  return OK;
}
mgp_error mgp_type_float(mgp_type **result) {
  *result = &g_type_float;
  return OK;
}
mgp_error mgp_type_node(mgp_type **result) {
  *result = &g_type_node;
  return OK;
}
mgp_error mgp_type_string(mgp_type **result) {  // :1692
  *result = &g_type_string;
  return OK;
}
mgp_error mgp_type_list(mgp_type *element, mgp_type **result) {  // :1741 (the three element types the modules here use)
  if (element == &g_type_node) *result = &g_type_list_node;
  else if (element == &g_type_float) *result = &g_type_list_float;
  else if (element == &g_type_int) *result = &g_type_list_int;
  else return mgp_error::MGP_ERROR_NOT_YET_IMPLEMENTED;
  return OK;
}
mgp_error mgp_module_add_read_procedure(mgp_module *module, const char *name, mgp_proc_cb cb, mgp_proc **result) {
  if (!ValidIdentifier(name)) return mgp_error::MGP_ERROR_INVALID_ARGUMENT;
  if (module->procs.count(name)) return mgp_error::MGP_ERROR_LOGIC_ERROR;
  mgp_proc &p = module->procs[name];
  p.name = name;
  p.cb = cb;
  *result = &p;
  return OK;
}
mgp_error mgp_proc_add_arg(mgp_proc *proc, const char *name, mgp_type *type) {  // :4880-4920
  if (!ValidIdentifier(name)) return mgp_error::MGP_ERROR_INVALID_ARGUMENT;
  for (auto &a : proc->opt_args)
    if (!a.required) return mgp_error::MGP_ERROR_LOGIC_ERROR;  // required args must precede optional ones
  ProcArg a;
  a.name = name;
  a.type = type;
  a.required = true;
  proc->opt_args.push_back(a);
  return OK;
}
mgp_error mgp_value_get_vertex(mgp_value *val, mgp_vertex **result) {  // unchecked union read in the real host
  *result = val->vertex;
  return OK;
}
mgp_error mgp_proc_add_opt_arg(mgp_proc *proc, const char *name, mgp_type *type, mgp_value *default_value) {
  if (!ValidIdentifier(name)) return mgp_error::MGP_ERROR_INVALID_ARGUMENT;
  if (default_value->kind == ValueKind::Vertex) return mgp_error::MGP_ERROR_VALUE_CONVERSION;
  if (default_value->kind != type->kind) return mgp_error::MGP_ERROR_LOGIC_ERROR;
  ProcArg a;
  a.name = name;
  a.type = type;
  a.default_value.kind = default_value->kind;  // copied; the caller still destroys its value
  a.default_value.i = default_value->i;
  a.default_value.d = default_value->d;
  a.default_value.s = default_value->s;
  proc->opt_args.push_back(a);
  return OK;
}
mgp_error mgp_proc_add_result(mgp_proc *proc, const char *name, mgp_type *type) {
  if (!ValidIdentifier(name)) return mgp_error::MGP_ERROR_INVALID_ARGUMENT;
  for (auto &r : proc->results)
    if (r.first == name) return mgp_error::MGP_ERROR_LOGIC_ERROR;
  proc->results.emplace_back(name, type);
  return OK;
}
mgp_error mgp_log(mgp_log_level, const char *) { return OK; }
int mgp_must_abort(mgp_graph *graph) {
  graph->data->abort_polls++;
  return graph->data->abort_flag.load();
}

// ---- test-driver API (fh_*) -------------------------------------------------------------------------------

struct FhModule {
  void *dl = nullptr;
  mgp_module registry;
  std::string name;
  int (*shutdown)() = nullptr;
};

struct FhResult {
  std::string error;  // empty = success
  std::vector<int64_t> nodes;
  std::vector<double> ranks;
};

// Vertices are iterated in ascending gid order (in-memory storage); `gids` may come in any order.
void *fh_graph_create(uint64_t n, const int64_t *gids, uint64_t m, const int64_t *src_gid, const int64_t *dst_gid,
                      int transactional) {
  auto *d = new FakeGraphData();
  d->gids.assign(gids, gids + n);
  std::sort(d->gids.begin(), d->gids.end());
  d->out.resize(n);
  d->hidden.assign(n, 0);
  for (uint32_t i = 0; i < n; ++i) d->index_of[d->gids[i]] = i;
  for (uint64_t e = 0; e < m; ++e) {
    auto s = d->index_of.find(src_gid[e]);
    auto t = d->index_of.find(dst_gid[e]);
    if (s == d->index_of.end() || t == d->index_of.end()) {
      delete d;
      return nullptr;
    }
    d->out[s->second].push_back(t->second);
  }
  d->edge_count = m;
  d->transactional = transactional;
  return d;
}
// One numeric property on the edges, in the order the edges were given to fh_graph_create (NaN = the edge lacks it).
int fh_graph_set_edge_property(void *graph, const char *name, uint64_t m, const int64_t *src_gid, const double *values,
                               int as_int) {
  auto *d = static_cast<FakeGraphData *>(graph);
  if (m != d->edge_count) return 1;
  d->prop_name = name;
  d->prop_as_int = as_int != 0;
  d->out_prop.assign(d->out.size(), {});
  for (uint64_t e = 0; e < m; ++e) d->out_prop[d->index_of.at(src_gid[e])].push_back(values[e]);  // same insertion order as `out`
  for (size_t v = 0; v < d->out.size(); ++v)
    if (d->out_prop[v].size() != d->out[v].size()) return 1;
  return 0;
}
void fh_graph_destroy(void *graph) { delete static_cast<FakeGraphData *>(graph); }
void fh_graph_set_abort(void *graph, int flag) { static_cast<FakeGraphData *>(graph)->abort_flag = flag; }
long fh_graph_abort_polls(void *graph) { return static_cast<FakeGraphData *>(graph)->abort_polls.load(); }
int fh_graph_hide_vertex(void *graph, int64_t gid) {
  auto *d = static_cast<FakeGraphData *>(graph);
  auto it = d->index_of.find(gid);
  if (it == d->index_of.end()) return 1;
  d->hidden[it->second] = 1;
  return 0;
}
long fh_live_objects(void) { return g_live_objects.load(); }

// SharedLibraryModule::Load (module.cpp:855-918)
void *fh_module_load(const char *path, char *err, size_t err_cap) {
  auto fail = [&](const std::string &msg) -> void * {
    if (err && err_cap) {
      std::strncpy(err, msg.c_str(), err_cap - 1);
      err[err_cap - 1] = '\0';
    }
    return nullptr;
  };
  dlerror();
  void *dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!dl) return fail(std::string("Unable to load module; ") + dlerror());
  auto init = reinterpret_cast<int (*)(mgp_module *, mgp_memory *)>(dlsym(dl, "mgp_init_module"));
  if (!init) {
    dlclose(dl);
    return fail("Unable to load module; mgp_init_module not found");
  }
  auto *m = new FhModule();
  m->dl = dl;
  std::string p(path);
  const size_t slash = p.find_last_of('/');
  std::string file = slash == std::string::npos ? p : p.substr(slash + 1);
  const size_t dot = file.find('.');
  m->name = dot == std::string::npos ? file : file.substr(0, dot);  // module name = file stem (module.cpp:1430-1448)
  mgp_memory memory;
  const int rc = init(&m->registry, &memory);
  if (rc != 0) {
    dlclose(dl);
    delete m;
    return fail("Unable to load module; mgp_init_module_failed: " + std::to_string(rc));
  }
  m->shutdown = reinterpret_cast<int (*)()>(dlsym(dl, "mgp_shutdown_module"));
  return m;
}
int fh_module_close(void *module) {
  auto *m = static_cast<FhModule *>(module);
  int rc = 0;
  if (m->shutdown) rc = m->shutdown();
  dlclose(m->dl);
  delete m;
  return rc;
}
// "name(arg = default :: TYPE, ...) :: (field :: TYPE, ...)" like the engine prints signatures
int fh_module_signature(void *module, const char *proc_name, char *buf, size_t cap) {
  auto *m = static_cast<FhModule *>(module);
  auto it = m->registry.procs.find(proc_name);
  if (it == m->registry.procs.end()) return 1;
  std::string s = it->second.name + "(";
  bool first = true;
  for (auto &a : it->second.opt_args) {
    if (!first) s += ", ";
    first = false;
    if (a.required) {
      s += a.name + " :: " + a.type->name;
      continue;
    }
    char num[64];
    if (a.default_value.kind == ValueKind::String) {
      s += a.name + " = \"" + a.default_value.s + "\" :: " + a.type->name;
      continue;
    }
    if (a.default_value.kind == ValueKind::Int)
      snprintf(num, sizeof(num), "%lld", static_cast<long long>(a.default_value.i));
    else
      snprintf(num, sizeof(num), "%.17g", a.default_value.d);
    s += a.name + " = " + num + " :: " + a.type->name;
  }
  s += ") :: (";
  first = true;
  for (auto &r : it->second.results) {
    if (!first) s += ", ";
    first = false;
    s += r.first + " :: " + r.second->name;
  }
  s += ")";
  if (buf && cap) {
    std::strncpy(buf, s.c_str(), cap - 1);
    buf[cap - 1] = '\0';
  }
  return 0;
}

// CallCustomProcedure (operator.cpp:7703-7800) + argument validation (module.cpp:1584-1601,
// module.hpp:169-202).  kinds[i] is 'i' (integer literal) or 'd' (float literal).
// kinds: 'i' INTEGER, 'd' FLOAT, 'v' NODE (gid in ivals), 's' STRING (svals[i]), 'V' LIST OF NODE, 'D' LIST OF FLOAT;
// a list argument i takes list_len[i] consecutive entries of list_ivals (gids) / list_dvals, in argument order.
void *fh_call_ex(void *module, const char *proc_name, void *graph_data, int n_args, const char *kinds,
                 const int64_t *ivals, const double *dvals, const char *const *svals, const uint64_t *list_len,
                 const int64_t *list_ivals, const double *list_dvals) {
  auto *m = static_cast<FhModule *>(module);
  auto *res_out = new FhResult();
  auto it = m->registry.procs.find(proc_name);
  if (it == m->registry.procs.end()) {
    res_out->error = "There is no procedure named '" + m->name + "." + proc_name + "'.";
    return res_out;
  }
  mgp_proc &proc = it->second;
  const std::string fq = m->name + "." + proc.name;
  size_t n_required = 0;
  for (auto &a : proc.opt_args) n_required += a.required;
  if (n_args < 0 || static_cast<size_t>(n_args) > proc.opt_args.size() || static_cast<size_t>(n_args) < n_required) {
    res_out->error = "'" + fq + "' requires between " + std::to_string(n_required) + " and " +
                     std::to_string(proc.opt_args.size()) + " arguments.";
    return res_out;
  }
  mgp_graph graph{static_cast<FakeGraphData *>(graph_data)};
  std::vector<std::unique_ptr<mgp_vertex>> arg_vertices;
  uint64_t list_cursor = 0;
  mgp_list args;
  for (size_t i = 0; i < proc.opt_args.size(); ++i) {
    mgp_value v;
    if (i < static_cast<size_t>(n_args)) {
      const char k = kinds[i];
      const ValueKind given = k == 'i' ? ValueKind::Int : k == 'v' ? ValueKind::Vertex : k == 's' ? ValueKind::String
                              : (k == 'V' || k == 'D') ? ValueKind::List : ValueKind::Double;
      const mgp_type *want = proc.opt_args[i].type;
      const bool element_ok = given != ValueKind::List || (want->element && want->element->kind == (k == 'V' ? ValueKind::Vertex : ValueKind::Double));
      if (given != want->kind || !element_ok) {  // strict: an integer literal does not satisfy FLOAT
        res_out->error = "'" + fq + "' argument named '" + proc.opt_args[i].name + "' at position " +
                         std::to_string(i) + " must be of type " + proc.opt_args[i].type->name + ".";
        return res_out;
      }
      v.kind = given;
      v.i = ivals[i];
      v.d = dvals[i];
      if (given == ValueKind::String) v.s = (svals && svals[i]) ? svals[i] : "";
      if (given == ValueKind::List) {
        v.list = std::make_shared<mgp_list>();
        for (uint64_t j = 0; j < (list_len ? list_len[i] : 0); ++j, ++list_cursor) {
          mgp_value item;
          if (k == 'V') {
            auto found = graph.data->index_of.find(list_ivals[list_cursor]);
            if (found == graph.data->index_of.end()) {
              res_out->error = "'" + fq + "': no vertex with the given id";
              return res_out;
            }
            arg_vertices.push_back(std::make_unique<mgp_vertex>(mgp_vertex{&graph, found->second, false}));
            item.kind = ValueKind::Vertex;
            item.vertex = arg_vertices.back().get();
          } else {
            item.kind = ValueKind::Double;
            item.d = list_dvals[list_cursor];
          }
          v.list->items.push_back(item);
        }
      }
      if (given == ValueKind::Vertex) {  // a NODE argument is passed by gid
        auto found = graph.data->index_of.find(ivals[i]);
        if (found == graph.data->index_of.end()) {
          res_out->error = "'" + fq + "': no vertex with the given id";
          return res_out;
        }
        arg_vertices.push_back(std::make_unique<mgp_vertex>(mgp_vertex{&graph, found->second, false}));
        v.vertex = arg_vertices.back().get();
      }
    } else {
      v = proc.opt_args[i].default_value;
    }
    args.items.push_back(v);
  }
  mgp_result result;
  result.proc = &proc;
  mgp_memory memory;
  proc.cb(&args, &graph, &result, &memory);
  if (result.has_error) {
    res_out->error = fq + ": " + result.error;
    return res_out;
  }
  for (auto &row : result.rows) {
    int64_t node = -1;
    double rank = 0.0;
    for (size_t f = 0; f < proc.results.size(); ++f) {
      if (proc.results[f].first == "node") node = row->fields[f].i;
      if (proc.results[f].first == "rank" || proc.results[f].first == "pagerank") rank = row->fields[f].d;
      if (proc.results[f].first == "distance") rank = static_cast<double>(row->fields[f].i);  // INTEGER result column
    }
    res_out->nodes.push_back(node);
    res_out->ranks.push_back(rank);
  }
  return res_out;
}
void *fh_call(void *module, const char *proc_name, void *graph_data, int n_args, const char *kinds,
              const int64_t *ivals, const double *dvals) {
  return fh_call_ex(module, proc_name, graph_data, n_args, kinds, ivals, dvals, nullptr, nullptr, nullptr, nullptr);
}

const char *fh_result_error(void *result) {
  auto *r = static_cast<FhResult *>(result);
  return r->error.empty() ? nullptr : r->error.c_str();
}
uint64_t fh_result_rows(void *result) { return static_cast<FhResult *>(result)->nodes.size(); }
void fh_result_copy(void *result, int64_t *nodes, double *ranks) {
  auto *r = static_cast<FhResult *>(result);
  if (!r->nodes.empty()) {
    std::memcpy(nodes, r->nodes.data(), r->nodes.size() * sizeof(int64_t));
    std::memcpy(ranks, r->ranks.data(), r->ranks.size() * sizeof(double));
  }
}
void fh_result_destroy(void *result) { delete static_cast<FhResult *>(result); }

}  // extern "C"
