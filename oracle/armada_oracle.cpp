// armada_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// Single-threaded C++17 restatement of the Armada scheduling round.  Every function
// cites the reference file:line (relative to /root/reference/internal/scheduler/) that it
// follows.  Data structures are the natural C++ equivalents of the reference's:
//   go-memdb ordered index  -> std::set per (priority level, node type)
//   memdb write txn         -> undo log (abort = replay backwards)
//   container/heap PQs      -> linear arg-min (both `Less` are strict total orders, so any
//                              correct priority queue pops in the same order)
// Parity: pinned by the reference's own table tests transcribed under tests/ (see header).
#include "armada_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace {

constexpr uint32_t NONE = ARMADA_NONE;
constexpr int32_t EVICTED_PRIORITY = -1;  // internaltypes/node.go:19
constexpr int32_t MIN_PRIORITY = -1;      // internaltypes/node.go:21
constexpr int32_t NON_PREEMPTIBLE_CUTOFF = std::numeric_limits<int32_t>::max();  // nodedb.go:1023

thread_local std::string g_err;

struct Fail {
  int32_t code;
  std::string msg;
};
[[noreturn]] void fail(int32_t code, const std::string& m) { throw Fail{code, m}; }

// ---------------------------------------------------------------------------------------
// Ordered node index (nodedb/encoding.go:37-58): key = (rounded indexed resources..., nodeIndex)
// ---------------------------------------------------------------------------------------
struct IndexKey {
  int64_t k[ARMADA_MAX_RESOURCES];
  uint64_t node_index;
  uint32_t node;
};
struct KeyLess {
  int r;
  bool operator()(const IndexKey& a, const IndexKey& b) const {
    for (int i = 0; i < r; ++i) {
      if (a.k[i] != b.k[i]) return a.k[i] < b.k[i];
    }
    return a.node_index < b.node_index;
  }
};
using IndexSet = std::set<IndexKey, KeyLess>;

// roundQuantityToResolution, nodedb/encoding.go:56-58 (Go integer division truncates toward 0)
inline int64_t round_to_resolution(int64_t q, int64_t res) { return (q / res) * res; }

// JobSchedulingContext + PodSchedulingContext (scheduling/context/job.go:23-70, pod.go:36-57)
struct JobCtx {
  uint32_t job = NONE;
  bool is_evicted = false;
  uint32_t assigned_node = NONE;   // AssignedNode (+ nodeIdLabel selector)
  uint32_t row = NONE;             // static bitmap row incl. AdditionalTolerations
  bool has_additional = false;     // AdditionalNodeSelectors/Tolerations non-empty
  int current_gang_cardinality = 1;
  uint8_t reason = ARMADA_REASON_NONE;  // UnschedulableReason ("" == NONE)
  bool has_pctx = false;
  uint32_t p_node = NONE;
  int32_t p_sched_at = 0;
  int32_t p_preempted_at = MIN_PRIORITY;
  uint8_t p_method = ARMADA_METHOD_NONE;
  bool p_away = false;
  // PodSchedulingContext.NumExcludedNodesByReason (context/pod.go:51) by reason KIND (ARMADA_EXCL_*)
  uint32_t excl[ARMADA_EXCLUDED_KINDS] = {0, 0, 0, 0, 0};
  bool is_successful() const { return reason == ARMADA_REASON_NONE; }  // job.go:113-115
  bool pctx_successful() const { return has_pctx && p_node != NONE; }  // pod.go:59-61
};

struct Undo {
  enum Kind { ALLOC, BIND, REBIND, UNBIND, EVTABLE_DEL } kind;
  uint32_t a = 0, b = 0;
  int idx = 0;
  bool flag = false;
  std::vector<int64_t> old;
};

// ---------------------------------------------------------------------------------------
// NodeDb (nodedb/nodedb.go)
// ---------------------------------------------------------------------------------------
struct NodeDb {
  const ArmadaRoundInput* in;
  int D, R, PL, T, PC;
  uint32_t N, J;
  std::vector<int64_t> alloc;            // [PL][D][N]  Node.AllocatableByPriority
  std::vector<int64_t> cur_key;          // [N][PL][R]  Node.Keys
  std::vector<IndexSet> index;           // [PL][T]     memdb index per priority (+type prefix)
  std::vector<uint32_t> nodes_of_type;   // [T]         numNodesByNodeType (nodedb.go:1148-1159)
  std::vector<std::vector<uint32_t>> node_jobs;  // keys of Node.AllocatedByJobId
  std::vector<uint32_t> bound_node;      // [J] node whose AllocatedByJobId holds the job
  std::vector<uint8_t> evicted_on_node;  // [J] Node.EvictedJobRunIds
  std::vector<int32_t> sched_prio;       // scheduledAtPriorityByJobId (nodedb.go:140-148)
  std::vector<uint8_t> has_sched_prio;
  std::map<int, uint32_t> evicted_by_index;  // "evictedJobs" table, index "index"
  // ---- exact acceleration of selectNodeForJobWithFairPreemption (select_fair_preemption_indexed) ----
  // every node's evicted-table entries (descending walk = reverse iteration), the node each entry sits
  // on, a log of the nodes whose rows or entries changed, and per (job class, static row) every node's
  // "trigger index" in an ordered set.  Test infrastructure like the rest of the oracle; checked
  // against the literal walk by tests/test_golden_units.py (ARMADA_ORACLE_FAIR=check).
  std::vector<std::set<int>> ev_on_node;
  std::map<int, uint32_t> ev_node_of_index;
  std::vector<uint32_t> dirty_log;
  struct FairIdx {
    size_t cursor = 0;
    std::vector<int> trig;
    std::set<std::pair<int, uint32_t>> order;
  };
  std::map<std::pair<uint32_t, uint32_t>, FairIdx> fair_cache;
  int fair_mode = 0;  // 0 literal below 4096 entries / indexed above, 1 literal, 2 indexed, 3 both + compare
  void ev_insert(int idx, uint32_t job) {
    evicted_by_index[idx] = job;
    const uint32_t n = bound_node[job];
    if (n != NONE) {
      if (ev_on_node.size() != N) ev_on_node.assign(N, {});
      ev_on_node[n].insert(idx);
      ev_node_of_index[idx] = n;
      dirty_log.push_back(n);
    }
  }
  void ev_erase(int idx) {
    evicted_by_index.erase(idx);
    auto it = ev_node_of_index.find(idx);
    if (it != ev_node_of_index.end()) {
      ev_on_node[it->second].erase(idx);
      dirty_log.push_back(it->second);
      ev_node_of_index.erase(it);
    }
  }
  void ev_clear() {
    evicted_by_index.clear();
    ev_node_of_index.clear();
    for (auto& s : ev_on_node) s.clear();
    fair_cache.clear();
    dirty_log.clear();
  }
  std::vector<int> evicted_index_of_job;     // "evictedJobs" table, index "id" (-1 = absent)
  std::vector<uint32_t> nodes_by_id;         // node indices in id order ("nodes" table index "id")
  bool in_txn = false;
  std::vector<Undo> undo;
  uint64_t stat_probes = 0, stat_fair_scans = 0;

  explicit NodeDb(const ArmadaRoundInput* input) : in(input) {
    if (const char* e = getenv("ARMADA_ORACLE_FAIR")) {
      std::string m(e);
      fair_mode = m == "literal" ? 1 : m == "indexed" ? 2 : m == "check" ? 3 : 0;
    }
    D = (int)in->num_resources;
    R = (int)in->num_indexed;
    PL = (int)in->num_priorities;
    T = (int)in->num_node_types;
    PC = (int)in->num_priority_classes;
    N = in->num_nodes;
    J = in->num_jobs;
    alloc.assign((size_t)PL * D * N, 0);
    cur_key.assign((size_t)N * PL * R, 0);
    index.assign((size_t)PL * T, IndexSet(KeyLess{R}));
    node_jobs.resize(N);
    nodes_of_type.assign((size_t)T, 0);
    for (uint32_t n = 0; n < N; ++n) nodes_of_type[in->node_type[n]]++;
    bound_node.assign(J, NONE);
    evicted_on_node.assign(J, 0);
    sched_prio.assign(J, 0);
    has_sched_prio.assign(J, 0);
    evicted_index_of_job.assign(J, -1);
    nodes_by_id.resize(N);
    for (uint32_t n = 0; n < N; ++n) nodes_by_id[n] = n;
    std::sort(nodes_by_id.begin(), nodes_by_id.end(),
              [&](uint32_t a, uint32_t b) { return in->node_id_rank[a] < in->node_id_rank[b]; });
    // newAllocatableByPriorityAndResourceType, nodedb.go:1185-1191
    for (int p = 0; p < PL; ++p)
      for (int d = 0; d < D; ++d)
        for (uint32_t n = 0; n < N; ++n) A(p, d, n) = in->node_allocatable[(size_t)d * N + n];
    for (uint32_t n = 0; n < N; ++n) insert_keys(n);
    // KubernetesResourceRequirements = AllResourceRequirements without the floating resources (job.go):
    // what the NodeDb compares with and books on nodes
    node_req.assign(in->class_request, in->class_request + (size_t)in->num_classes * D);
    for (uint32_t c = 0; c < in->num_classes; ++c)
      for (int d = 0; d < D; ++d)
        if ((in->floating_resource_mask >> d) & 1u) node_req[(size_t)c * D + d] = 0;
  }

  int64_t& A(int p, int d, uint32_t n) { return alloc[((size_t)p * D + d) * N + n]; }
  int64_t total(int d, uint32_t n) const { return in->node_total[(size_t)d * N + n]; }
  std::vector<int64_t> node_req;  // [C][D]
  const int64_t* req_of(uint32_t job) const { return node_req.data() + (size_t)in->job_class[job] * D; }
  // gang node uniformity (gang_scheduler.go:191): while >= 0, every jctx carries the node selector of this value
  // slot — the static rows come from class_uniformity_row instead of class_static_row / class_away_row
  int64_t uniformity_slot = -1;
  uint32_t base_row(uint32_t c) const {
    if (uniformity_slot < 0) return in->class_static_row[c];
    return in->class_uniformity_row[((size_t)c * uniformity_V() + (size_t)uniformity_slot) * (1 + ARMADA_MAX_AWAY)];
  }
  uint32_t away_row(uint32_t c, uint32_t k) const {
    if (uniformity_slot < 0) return in->class_away_row[(size_t)c * ARMADA_MAX_AWAY + k];
    return in->class_uniformity_row[((size_t)c * uniformity_V() + (size_t)uniformity_slot) * (1 + ARMADA_MAX_AWAY) + 1 + k];
  }
  size_t uniformity_V() const { return in->uniformity_value_start ? in->uniformity_value_start[in->num_uniformity_labels] : 0; }
  const ArmadaPriorityClass& pc_of(uint32_t job) const {
    return in->priority_classes[in->class_pc[in->job_class[job]]];
  }
  int level_of(int32_t priority) const {  // keyIndexByPriority, nodedb.go:1223-1234
    for (int p = 0; p < PL; ++p)
      if (in->priorities[p] == priority) return p;
    fail(ARMADA_E_INTERNAL, "no index for priority " + std::to_string(priority));
  }

  // nodeDbKey, nodedb.go:1279-1288
  IndexKey make_key(uint32_t n, int p) {
    IndexKey key{};
    for (int i = 0; i < R; ++i)
      key.k[i] = round_to_resolution(A(p, (int)in->indexed_resource[i], n), in->indexed_resolution[i]);
    key.node_index = in->node_index[n];
    key.node = n;
    return key;
  }
  void insert_keys(uint32_t n) {
    for (int p = 0; p < PL; ++p) {
      IndexKey key = make_key(n, p);
      for (int i = 0; i < R; ++i) cur_key[((size_t)n * PL + p) * R + i] = key.k[i];
      index[(size_t)p * T + in->node_type[n]].insert(key);
    }
  }
  // UpsertWithTxn, nodedb.go:1148-1159 (re-keys the node in every priority index)
  void upsert(uint32_t n) {
    dirty_log.push_back(n);
    for (int p = 0; p < PL; ++p) {
      IndexKey nk = make_key(n, p);
      int64_t* ck = &cur_key[((size_t)n * PL + p) * R];
      bool same = true;
      for (int i = 0; i < R; ++i) same = same && ck[i] == nk.k[i];
      if (same) continue;
      IndexKey ok = nk;
      for (int i = 0; i < R; ++i) ok.k[i] = ck[i];
      IndexSet& s = index[(size_t)p * T + in->node_type[n]];
      s.erase(ok);
      s.insert(nk);
      for (int i = 0; i < R; ++i) ck[i] = nk.k[i];
    }
  }

  // ---- txn (memdb write txn emulation) -------------------------------------------------
  void begin() {
    in_txn = true;
    undo.clear();
  }
  void commit() {
    in_txn = false;
    undo.clear();
  }
  void abort() {
    for (size_t i = undo.size(); i-- > 0;) {
      Undo& u = undo[i];
      switch (u.kind) {
        case Undo::ALLOC:
          for (int p = 0; p < PL; ++p)
            for (int d = 0; d < D; ++d) A(p, d, u.a) = u.old[(size_t)p * D + d];
          upsert(u.a);
          break;
        case Undo::BIND:  // job u.a had been added to node u.b
          remove_job(u.b, u.a);
          bound_node[u.a] = NONE;
          break;
        case Undo::REBIND:  // evicted flag of job u.a had been cleared
          evicted_on_node[u.a] = 1;
          break;
        case Undo::UNBIND:  // job u.a had been removed from node u.b
          node_jobs[u.b].push_back(u.a);
          bound_node[u.a] = u.b;
          evicted_on_node[u.a] = u.flag ? 1 : 0;
          break;
        case Undo::EVTABLE_DEL:
          ev_insert(u.idx, u.a);
          evicted_index_of_job[u.a] = u.idx;
          break;
      }
    }
    in_txn = false;
    undo.clear();
  }
  void log_alloc(uint32_t n) {
    if (!in_txn) return;
    Undo u;
    u.kind = Undo::ALLOC;
    u.a = n;
    u.old.resize((size_t)PL * D);
    for (int p = 0; p < PL; ++p)
      for (int d = 0; d < D; ++d) u.old[(size_t)p * D + d] = A(p, d, n);
    undo.push_back(std::move(u));
  }
  void remove_job(uint32_t n, uint32_t j) {
    auto& v = node_jobs[n];
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i] == j) {
        v[i] = v.back();
        v.pop_back();
        return;
      }
  }

  // markAllocatable, nodedb.go:1009-1019 (sign = -1 ⇒ markAllocated, :1005-1007)
  void mark(uint32_t n, int32_t cutoff, const int64_t* rs, int sign) {
    for (int p = 0; p < PL; ++p)
      if (in->priorities[p] <= cutoff)
        for (int d = 0; d < D; ++d) A(p, d, n) += sign * rs[d];
  }
  // priorityCutoffFor, nodedb.go:1031-1036
  int32_t cutoff_for(uint32_t job, int32_t scheduled_priority) const {
    return pc_of(job).preemptible ? scheduled_priority : NON_PREEMPTIBLE_CUTOFF;
  }

  // bindJobToNodeInPlace, nodedb.go:915-949
  void bind(uint32_t n, uint32_t job, int32_t priority) {
    log_alloc(n);
    bool is_evicted = bound_node[job] == n && evicted_on_node[job];
    if (is_evicted) {
      evicted_on_node[job] = 0;
      if (in_txn) {
        Undo u;
        u.kind = Undo::REBIND;
        u.a = job;
        undo.push_back(u);
      }
    } else {
      if (bound_node[job] != NONE)
        fail(ARMADA_E_INTERNAL, "job already has resources allocated on a node");
      node_jobs[n].push_back(job);
      bound_node[job] = n;
      if (in_txn) {
        Undo u;
        u.kind = Undo::BIND;
        u.a = job;
        u.b = n;
        undo.push_back(u);
      }
    }
    const int64_t* rq = req_of(job);
    mark(n, cutoff_for(job, priority), rq, -1);
    if (is_evicted) mark(n, EVICTED_PRIORITY, rq, +1);
    sched_prio[job] = priority;  // plain map: NOT part of the memdb txn (nodedb.go:946)
    has_sched_prio[job] = 1;
    upsert(n);
  }
  // evictJobFromNodeInPlace, nodedb.go:974-1003
  void evict(uint32_t job) {
    uint32_t n = bound_node[job];
    if (n == NONE) fail(ARMADA_E_INTERNAL, "job has no resources allocated on node");
    if (evicted_on_node[job]) fail(ARMADA_E_INTERNAL, "job is already evicted from node");
    if (!has_sched_prio[job]) fail(ARMADA_E_INTERNAL, "job not mapped to a priority");
    log_alloc(n);
    evicted_on_node[job] = 1;
    const int64_t* rq = req_of(job);
    mark(n, cutoff_for(job, sched_prio[job]), rq, +1);
    mark(n, EVICTED_PRIORITY, rq, -1);
    upsert(n);
  }
  // unbindJobFromNodeInPlace, nodedb.go:1059-1098
  void unbind(uint32_t job, uint32_t n) {
    bool is_evicted = bound_node[job] == n && evicted_on_node[job];
    if (bound_node[job] != n) {
      // "Job already unbound; nothing more to do." (EvictedJobRunIds entry deleted first)
      return;
    }
    log_alloc(n);
    if (in_txn) {
      Undo u;
      u.kind = Undo::UNBIND;
      u.a = job;
      u.b = n;
      u.flag = is_evicted;
      undo.push_back(u);
    }
    evicted_on_node[job] = 0;
    remove_job(n, job);
    bound_node[job] = NONE;
    const int64_t* rq = req_of(job);
    if (is_evicted) {
      mark(n, EVICTED_PRIORITY, rq, +1);
    } else {
      if (!has_sched_prio[job]) fail(ARMADA_E_INTERNAL, "job not mapped to a priority");
      mark(n, cutoff_for(job, sched_prio[job]), rq, +1);
    }
    upsert(n);
  }

  // CreateAndInsertWithJobDbJobsWithTxn, nodedb.go:43-60
  void bind_running_jobs() {
    for (uint32_t j = 0; j < J; ++j) {
      uint32_t n = in->job_node[j];
      if (n == NONE) continue;
      int32_t priority = in->job_scheduled_at_priority[j];
      if (priority == ARMADA_NO_PRIORITY) priority = pc_of(j).priority;
      bind(n, j, priority);
    }
  }

  // ---- matching (nodedb/nodematching.go) ----------------------------------------------
  bool row_bit(const uint32_t* bm, uint32_t row, uint32_t words, uint32_t i) const {
    return (bm[(size_t)row * words + (i >> 5)] >> (i & 31)) & 1u;
  }
  bool type_matches(uint32_t row, uint32_t t) const {  // NodeTypeJobRequirementsMet :127-139
    return row_bit(in->type_match, row, (in->num_node_types + 31) / 32, t);
  }
  // StaticJobRequirementsMet :161-190 (string predicates precomputed into static_match)
  bool static_met(uint32_t n, uint32_t row, const int64_t* rq) const {
    if (!row_bit(in->static_match, row, (in->num_static_classes + 31) / 32, in->node_static_class[n]))
      return false;
    for (int d = 0; d < D; ++d)
      if (rq[d] > total(d, n)) return false;
    return true;
  }
  // DynamicJobRequirementsMet :194-197 / resourceRequirementsMet :257-267
  bool dynamic_met_at(uint32_t n, int p, const int64_t* rq) {
    for (int d = 0; d < D; ++d)
      if (rq[d] > A(p, d, n)) return false;
    return true;
  }

  // ---- NodeTypeIterator (nodedb/nodeiteration.go:214-382) --------------------------------
  struct TypeIt {
    uint32_t type;
    IndexSet* set;
    IndexSet::iterator it;
    int64_t lower[ARMADA_MAX_RESOURCES];
    int64_t new_lower[ARMADA_MAX_RESOURCES];
  };
  void type_it_seek(TypeIt& ti) {  // newNodeTypeIterator :292-304 (LowerBound, nodeIndex 0)
    IndexKey lb{};
    for (int i = 0; i < R; ++i) lb.k[i] = ti.lower[i];
    lb.node_index = 0;
    ti.it = ti.set->lower_bound(lb);
  }
  // NextNode :318-382
  uint32_t type_it_next(TypeIt& ti, int p, const int64_t* ireq) {
    for (;;) {
      if (ti.it == ti.set->end()) return NONE;  // end of index, or node of another type
      uint32_t n = ti.it->node;
      ++ti.it;
      bool yielded = false, reseek = false;
      for (int i = 0; i < R; ++i) {
        int64_t node_q = A(p, (int)in->indexed_resource[i], n);
        ti.new_lower[i] = round_to_resolution(node_q, in->indexed_resolution[i]);
        if (node_q < ireq[i]) {
          for (int j = i; j < R; ++j) ti.new_lower[j] = ireq[j];
          // bytes.Compare(it.key, it.newKey) == -1
          bool less = false;
          for (int j = 0; j < R; ++j) {
            if (ti.lower[j] != ti.new_lower[j]) {
              less = ti.lower[j] < ti.new_lower[j];
              break;
            }
          }
          if (less) {
            for (int j = 0; j < R; ++j) std::swap(ti.lower[j], ti.new_lower[j]);
            reseek = true;
          }  // else: "new lower-bound is not greater than current bound": keep iterating
          break;
        } else if (i == R - 1) {
          yielded = true;
        }
      }
      if (yielded) return n;
      if (reseek) type_it_seek(ti);
    }
  }

  // ---- NodeTypesIterator (nodedb/nodeiteration.go:74-208): k-way merge over types ---------
  struct TypesIt {
    std::vector<TypeIt> its;
    std::vector<uint32_t> head;  // current head node per iterator (NONE = exhausted)
    int p;
    const int64_t* ireq;
  };
  // nodeTypesIteratorPQ.less :170-185 — UNROUNDED allocatable, then node id
  bool merge_less(uint32_t a, uint32_t b, int p) {
    for (int i = 0; i < R; ++i) {
      int64_t qa = A(p, (int)in->indexed_resource[i], a), qb = A(p, (int)in->indexed_resource[i], b);
      if (qa != qb) return qa < qb;
    }
    return in->node_id_rank[a] < in->node_id_rank[b];
  }
  void types_it_init(TypesIt& m, uint32_t row, int p, const int64_t* ireq) {
    m.p = p;
    m.ireq = ireq;
    for (uint32_t t = 0; t < (uint32_t)T; ++t) {
      if (!type_matches(row, t)) continue;  // NodeTypesMatchingJob, nodedb.go:1102-1117
      TypeIt ti;
      ti.type = t;
      ti.set = &index[(size_t)p * T + t];
      for (int i = 0; i < R; ++i) ti.lower[i] = ti.new_lower[i] = ireq[i];
      type_it_seek(ti);
      m.its.push_back(ti);
    }
    m.head.resize(m.its.size());
    for (size_t k = 0; k < m.its.size(); ++k) m.head[k] = type_it_next(m.its[k], p, ireq);
  }
  uint32_t types_it_next(TypesIt& m) {  // NextNode :133-149
    int best = -1;
    for (size_t k = 0; k < m.its.size(); ++k) {
      if (m.head[k] == NONE) continue;
      if (best < 0 || merge_less(m.head[k], m.head[(size_t)best], m.p)) best = (int)k;
    }
    if (best < 0) return NONE;
    uint32_t n = m.head[(size_t)best];
    m.head[(size_t)best] = type_it_next(m.its[(size_t)best], m.p, m.ireq);
    return n;
  }

  // selectNodeForPodAtPriority + selectNodeForPodWithItAtPriority, nodedb.go:717-805
  uint32_t select_at_priority(JobCtx& jc, int32_t priority) {
    ++stat_probes;
    const int64_t* rq = req_of(jc.job);
    int64_t ireq[ARMADA_MAX_RESOURCES];
    for (int i = 0; i < R; ++i) ireq[i] = rq[in->indexed_resource[i]];
    int p = level_of(priority);
    TypesIt m;
    types_it_init(m, jc.row, p, ireq);
    // pctx.NumExcludedNodesByReason = Clone(what NodeTypesMatchingJob excluded) (nodedb.go:605-640) + one
    // per node the walk reaches and JobRequirementsMet rejects (:786-797)
    for (uint32_t k = 0; k < ARMADA_EXCLUDED_KINDS; ++k) jc.excl[k] = 0;
    for (uint32_t t = 0; t < (uint32_t)T; ++t)
      if (!type_matches(jc.row, t)) jc.excl[ARMADA_EXCL_NODE_TYPE] += nodes_of_type[t];  // :1102-1117
    for (uint32_t n = types_it_next(m); n != NONE; n = types_it_next(m)) {
      // JobRequirementsMet, nodematching.go:147-157
      if (static_met(n, jc.row, rq) && dynamic_met_at(n, p, rq)) {
        jc.p_node = n;
        jc.p_preempted_at = priority;
        return n;
      }
      // StaticJobRequirementsMet checks taints, labels and affinity before the total resources (:161-190)
      const bool string_predicates = row_bit(in->static_match, jc.row, (in->num_static_classes + 31) / 32, in->node_static_class[n]);
      jc.excl[string_predicates ? ARMADA_EXCL_RESOURCES : ARMADA_EXCL_STATIC] += 1;
    }
    return NONE;
  }
  // the deferred function of SelectNodeForJobWithTxn (nodedb.go:449-462): a pod that found no node counts
  // the nodes nothing excluded explicitly as "insufficient resources"
  uint32_t fail_select(JobCtx& jc) {
    uint32_t expl = 0;
    for (uint32_t k = 0; k < ARMADA_EXCLUDED_KINDS; ++k) expl += jc.excl[k];
    if (N > expl) jc.excl[ARMADA_EXCL_IMPLICIT] += N - expl;
    return NONE;
  }

  // node n's trigger: the evicted-table index at which the reference's descending walk first finds
  // the node's running total (allocatable at the evicted priority + the evicted jobs seen so far)
  // large enough — and the node passes the static requirements; -1 = never selected
  int fair_trigger(uint32_t n, uint32_t row, const int64_t* rq) {
    if (n >= ev_on_node.size() || ev_on_node[n].empty()) return -1;
    int lvl_evicted = level_of(EVICTED_PRIORITY);
    int64_t avail[ARMADA_MAX_RESOURCES];
    for (int d = 0; d < D; ++d) avail[d] = A(lvl_evicted, d, n);
    for (auto it = ev_on_node[n].rbegin(); it != ev_on_node[n].rend(); ++it) {
      const int64_t* erq = req_of(evicted_by_index[*it]);
      bool dyn = true;
      for (int d = 0; d < D; ++d) {
        avail[d] += erq[d];
        if (rq[d] > avail[d]) dyn = false;
      }
      if (dyn) return static_met(n, row, rq) ? *it : -1;
    }
    return -1;
  }
  // The same selection as the literal walk below without re-walking the table: the walk reaches the
  // node with the LARGEST trigger index first.
  uint32_t select_fair_preemption_indexed(JobCtx& jc, bool commit) {
    const int64_t* rq = req_of(jc.job);
    auto key = std::make_pair(in->job_class[jc.job], jc.row);
    auto ins = fair_cache.try_emplace(key);
    FairIdx& F = ins.first->second;
    auto recompute = [&](uint32_t n) {
      int old = F.trig[n];
      int t = fair_trigger(n, jc.row, rq);
      if (old == t) return;
      if (old >= 0) F.order.erase({old, n});
      if (t >= 0) F.order.insert({t, n});
      F.trig[n] = t;
    };
    if (ins.second) {
      F.trig.assign(N, -1);
      for (uint32_t n = 0; n < (uint32_t)ev_on_node.size(); ++n)
        if (!ev_on_node[n].empty()) recompute(n);
      F.cursor = dirty_log.size();
    } else {
      for (; F.cursor < dirty_log.size(); ++F.cursor) recompute(dirty_log[F.cursor]);
    }
    if (F.order.empty()) return NONE;
    const int trig = F.order.rbegin()->first;
    const uint32_t n = F.order.rbegin()->second;
    if (!commit) return n;
    std::vector<std::pair<int, uint32_t>> victims;
    for (auto it = ev_on_node[n].rbegin(); it != ev_on_node[n].rend() && *it >= trig; ++it) victims.emplace_back(*it, evicted_by_index[*it]);
    int32_t max_priority = MIN_PRIORITY;
    for (auto& [eidx, pj] : victims) {
      unbind(pj, n);
      if (in_txn) {
        Undo u;
        u.kind = Undo::EVTABLE_DEL;
        u.a = pj;
        u.idx = eidx;
        undo.push_back(u);
      }
      ev_erase(eidx);
      evicted_index_of_job[pj] = -1;
      int32_t pr = has_sched_prio[pj] ? sched_prio[pj] : pc_of(pj).priority;
      if (pr > max_priority) max_priority = pr;
    }
    jc.p_node = n;
    jc.p_preempted_at = max_priority;
    return n;
  }

  // selectNodeForJobWithFairPreemption, nodedb.go:812-903
  uint32_t select_fair_preemption(JobCtx& jc) {
    ++stat_fair_scans;
    if (fair_mode == 2 || (fair_mode == 0 && evicted_by_index.size() >= 4096)) return select_fair_preemption_indexed(jc, true);
    if (fair_mode == 3) {
      const uint32_t want = select_fair_preemption_indexed(jc, false);
      const uint32_t got = select_fair_preemption_literal(jc);
      if (want != got) fail(ARMADA_E_INTERNAL, "indexed fair preemption disagrees with the literal walk");
      return got;
    }
    return select_fair_preemption_literal(jc);
  }
  uint32_t select_fair_preemption_literal(JobCtx& jc) {
    struct Considered {
      std::vector<int64_t> avail;
      std::vector<std::pair<int, uint32_t>> evicted;  // (index, job)
      bool static_not_met = false;
    };
    std::map<uint32_t, Considered> by_node;
    const int64_t* rq = req_of(jc.job);
    int32_t max_priority = MIN_PRIORITY;
    int lvl_evicted = level_of(EVICTED_PRIORITY);
    // ReverseLowerBound("evictedJobs", "index", MaxInt): descending index
    for (auto rit = evicted_by_index.rbegin(); rit != evicted_by_index.rend(); ++rit) {
      int idx = rit->first;
      uint32_t ej = rit->second;
      uint32_t n = bound_node[ej];
      if (n == NONE) fail(ARMADA_E_INTERNAL, "evicted job does not have an assigned nodeId");
      auto it = by_node.find(n);
      if (it == by_node.end()) {
        Considered c;
        c.avail.resize(D);
        for (int d = 0; d < D; ++d) c.avail[d] = A(lvl_evicted, d, n);
        it = by_node.emplace(n, std::move(c)).first;
      }
      Considered& c = it->second;
      if (c.static_not_met) continue;
      const int64_t* erq = req_of(ej);
      for (int d = 0; d < D; ++d) c.avail[d] += erq[d];
      c.evicted.emplace_back(idx, ej);
      bool dyn = true;
      for (int d = 0; d < D; ++d)
        if (rq[d] > c.avail[d]) dyn = false;
      if (!dyn) continue;
      if (!static_met(n, jc.row, rq)) {
        c.static_not_met = true;
        continue;
      }
      std::vector<std::pair<int, uint32_t>> victims = c.evicted;  // iteration ends here
      for (auto& [eidx, pj] : victims) {
        unbind(pj, n);
        if (in_txn) {
          Undo u;
          u.kind = Undo::EVTABLE_DEL;
          u.a = pj;
          u.idx = eidx;
          undo.push_back(u);
        }
        ev_erase(eidx);
        evicted_index_of_job[pj] = -1;
        int32_t pr = has_sched_prio[pj] ? sched_prio[pj] : pc_of(pj).priority;
        if (pr > max_priority) max_priority = pr;
      }
      jc.p_node = n;
      jc.p_preempted_at = max_priority;
      return n;
    }
    return NONE;
  }

  // selectNodeForJobWithUrgencyPreemption, nodedb.go:682-715
  uint32_t select_urgency(JobCtx& jc) {
    for (int p = 0; p < PL; ++p) {
      int32_t priority = in->priorities[p];
      if (priority == EVICTED_PRIORITY) continue;
      if (priority > jc.p_sched_at) break;
      uint32_t n = select_at_priority(jc, priority);
      if (n != NONE) return n;
    }
    return NONE;
  }

  // selectNodeForJobWithTxnAtPriority, nodedb.go:605-666
  uint32_t select_home_or_away_at_priority(JobCtx& jc) {
    uint32_t n = select_at_priority(jc, EVICTED_PRIORITY);
    if (n != NONE) {
      jc.p_method = ARMADA_METHOD_NO_PREEMPTION;
      return n;
    }
    n = select_at_priority(jc, jc.p_sched_at);
    if (n == NONE) return NONE;
    jc.p_node = NONE;
    jc.p_preempted_at = MIN_PRIORITY;
    n = select_fair_preemption(jc);
    if (n != NONE) {
      jc.p_method = ARMADA_METHOD_FAIRSHARE;
      return n;
    }
    jc.p_node = NONE;
    jc.p_preempted_at = MIN_PRIORITY;
    n = select_urgency(jc);
    if (n != NONE) {
      jc.p_method = ARMADA_METHOD_URGENCY;
      return n;
    }
    return NONE;
  }

  // SelectNodeForJobWithTxn, nodedb.go:431-512
  uint32_t select_node_for_job(JobCtx& jc) {
    const ArmadaPriorityClass& pc = pc_of(jc.job);
    int32_t priority = has_sched_prio[jc.job] ? sched_prio[jc.job] : pc.priority;
    jc.has_pctx = true;
    jc.p_node = NONE;
    jc.p_sched_at = priority;
    jc.p_preempted_at = MIN_PRIORITY;
    jc.p_method = ARMADA_METHOD_NONE;
    jc.p_away = false;
    for (uint32_t k = 0; k < ARMADA_EXCLUDED_KINDS; ++k) jc.excl[k] = 0;
    const int64_t* rq = req_of(jc.job);
    if (jc.assigned_node != NONE) {  // :465-476 + selectNodeForPodWithItAtPriority(onlyDynamic)
      uint32_t n = jc.assigned_node;
      uint8_t fl = in->node_flags[n];
      bool matches;
      if ((fl & ARMADA_NODE_UNSCHEDULABLE) && (fl & ARMADA_NODE_OVERALLOCATED)) {
        matches = true;  // :774-781
      } else {
        matches = dynamic_met_at(n, level_of(priority), rq);
      }
      jc.p_method = ARMADA_METHOD_RESCHEDULED;
      if (matches) {
        jc.p_node = n;
        jc.p_preempted_at = priority;
        return n;
      }
      jc.excl[ARMADA_EXCL_RESOURCES] = 1;  // selectNodeForPodWithItAtPriority on the one node (:786-797)
      return fail_select(jc);
    }
    for (int d = 0; d < D; ++d)  // :478-483
      if (((in->disallowed_resource_mask >> d) & 1u) && rq[d] > 0) {
        jc.excl[ARMADA_EXCL_DISALLOWED] = N;
        return fail_select(jc);
      }
    if (!in->disable_home_scheduling) {
      uint32_t n = select_home_or_away_at_priority(jc);
      if (n != NONE) return n;
    }
    bool in_gang = in->job_gang[jc.job] != NONE;
    bool away_disabled = in->disable_away_scheduling || (in_gang && in->disable_gang_away_scheduling);
    if (!away_disabled) {
      uint32_t c = in->job_class[jc.job];
      for (uint32_t k = 0; k < pc.num_away; ++k) {
        // selectNodeForJobWithTxnAndAwayNodeType, nodedb.go:559-603
        uint32_t away_row = this->away_row(c, k);
        if (away_row == NONE) continue;  // "No extra taints to tolerate"
        uint32_t saved_row = jc.row;
        bool saved_additional = jc.has_additional;
        jc.row = away_row;  // AdditionalTolerations += tolerations for the away taints
        jc.has_additional = true;
        jc.p_sched_at = pc.away_priority[k];
        uint32_t n = select_home_or_away_at_priority(jc);
        if (n != NONE) {
          jc.p_method = ARMADA_METHOD_AWAY;
          jc.p_away = true;
          return n;
        }
        jc.row = saved_row;  // restore the tolerations slice
        jc.has_additional = saved_additional;
      }
    }
    return fail_select(jc);
  }

  // ScheduleManyWithTxn, nodedb.go:386-418
  bool schedule_many(std::vector<JobCtx*>& gang) {
    for (JobCtx* jc : gang) {
      jc->reason = ARMADA_REASON_NONE;
      uint32_t n = select_node_for_job(*jc);
      if (n == NONE) return false;
      bind(n, jc->job, jc->p_sched_at);
      // deleteEvictedJobSchedulingContextIfExistsWithTxn
      int idx = evicted_index_of_job[jc->job];
      if (idx >= 0) {
        if (in_txn) {
          Undo u;
          u.kind = Undo::EVTABLE_DEL;
          u.a = jc->job;
          u.idx = idx;
          undo.push_back(u);
        }
        ev_erase(idx);
        evicted_index_of_job[jc->job] = -1;
      }
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------
// DominantResourceFairness (scheduling/fairness/fairness.go:99-105) with
// DivideZeroOnError / Multiply / Max (internaltypes/resource_list.go:265-279,
// resource_fraction_list.go:17-39).
// ---------------------------------------------------------------------------------------
double drf_unweighted(int D, const int64_t* total, const double* mult, const int64_t* a) {
  double result = -std::numeric_limits<double>::infinity();
  for (int d = 0; d < D; ++d) {
    double frac = 0.0;
    if (total[d] != 0) frac = (double)a[d] / (double)total[d];
    double v = frac * mult[d];
    if (v > result) result = v;
  }
  // Go builtin max(0, x): NaN propagates, +0 preferred over -0.
  if (std::isnan(result)) return result;
  return result > 0.0 ? result : 0.0;
}

// ---------------------------------------------------------------------------------------
// The round
// ---------------------------------------------------------------------------------------
struct Round {
  const ArmadaRoundInput* in;
  NodeDb db;
  int D, PC;
  uint32_t Q, J, C;

  // QueueSchedulingContext (scheduling/context/queue.go:21-95)
  struct QueueCtx {
    double weight;
    std::vector<int64_t> allocated;        // [D]
    std::vector<int64_t> allocated_by_pc;  // [PC][D]
    std::vector<int64_t> penalty;          // [D] ShortJobPenalty
    double fair_share = 0, demand_capped = 0, uncapped = 0;
    double tokens = 0;
    int64_t burst = 0;
    bool inf = false;
  };
  std::vector<QueueCtx> qctx;
  // per-job membership of the qctx maps (queue.go:83-93)
  std::vector<uint8_t> successful, unsuccessful, rescheduled, evicted_by_id;
  std::vector<uint8_t> unsuccessful_reason;
  // SchedulingContext (scheduling/context/scheduling.go:27-70)
  std::vector<int64_t> scheduled_resources, evicted_resources;
  int64_t num_scheduled_jobs = 0, num_scheduled_gangs = 0, num_evicted_jobs = 0;
  uint32_t termination_reason = ARMADA_REASON_NONE;
  double weight_sum = 0;
  double global_tokens;
  // UnfeasibleSchedulingKeys: class -> reason of the first failed jctx (0 = absent)
  std::vector<uint8_t> unfeasible_key;
  size_t num_unfeasible = 0;
  // latest pod scheduling result per job (the jctx the result lists would carry)
  std::vector<JobCtx> last_jctx;
  std::vector<uint32_t> job_seq;  // loop iteration of the job's last gang attempt
  // the first schedule pass as the loop saw it (what QueueStats are made of, queue_scheduler.go:190-235)
  bool in_first_pass = false;
  std::vector<uint32_t> job_seq_first;
  std::vector<uint8_t> job_reason_first;
  std::vector<std::vector<uint32_t>> queued_by_queue;
  ArmadaRoundStats stats{};
  std::vector<std::vector<uint32_t>> gang_members;  // jobRepo.GetGangJobsByGangId

  explicit Round(const ArmadaRoundInput* input) : in(input), db(input) {
    D = (int)in->num_resources;
    PC = (int)in->num_priority_classes;
    Q = in->num_queues;
    J = in->num_jobs;
    C = in->num_classes;
    successful.assign(J, 0);
    unsuccessful.assign(J, 0);
    rescheduled.assign(J, 0);
    evicted_by_id.assign(J, 0);
    unsuccessful_reason.assign(J, 0);
    scheduled_resources.assign(D, 0);
    evicted_resources.assign(D, 0);
    unfeasible_key.assign(C, 0);
    last_jctx.resize(J);
    job_seq.assign(J, 0);
    job_seq_first.assign(J, 0);
    job_reason_first.assign(J, 0);
    global_tokens = in->global_limiter_tokens;
    gang_members.resize(in->num_gangs);
    for (uint32_t j = 0; j < J; ++j)
      if (in->job_gang[j] != NONE) gang_members[in->job_gang[j]].push_back(j);
    build_queue_contexts();
    build_queued_order();
  }

  const int64_t* req_of(uint32_t job) const { return in->class_request + (size_t)in->job_class[job] * D; }  // AllResourceRequirements
  uint32_t pc_index(uint32_t job) const { return in->class_pc[in->job_class[job]]; }

  // Snapshot construction when the caller leaves the queue accounting to the library (NULL
  // queue_allocated_by_pc / queue_constrained_demand): calculateJobSchedulingInfo
  // (scheduling_algo.go:522-632: allocation = running jobs per queue and priority class, demand = every
  // job of the pool, running jobs only for a cordoned queue) and constructSchedulingContext (:664-676:
  // constrained demand = the per-priority-class demand capped at the per-queue limit,
  // constraints.go:187-197, summed over the classes).
  std::vector<int64_t> der_alloc_pc, der_cdemand;
  void derive_snapshot() {
    der_alloc_pc.assign((size_t)Q * PC * D, 0);
    std::vector<int64_t> demand_pc((size_t)Q * PC * D, 0);
    for (uint32_t j = 0; j < J; ++j) {
      uint32_t q = in->job_queue[j];
      if (q == NONE) continue;
      const uint32_t cls = in->job_class[j];
      const uint32_t pc = in->class_pc[cls];
      const bool running = in->job_node[j] != NONE;
      const bool cordoned = in->queue_cordoned && in->queue_cordoned[q];
      for (int d = 0; d < D; ++d) {
        const int64_t r = in->class_request[(size_t)cls * D + d];
        if (running) der_alloc_pc[((size_t)q * PC + pc) * D + d] += r;
        if (running || !cordoned) demand_pc[((size_t)q * PC + pc) * D + d] += r;
      }
    }
    der_cdemand.assign((size_t)Q * D, 0);
    for (uint32_t q = 0; q < Q; ++q)
      for (int pc = 0; pc < PC; ++pc)
        for (int d = 0; d < D; ++d) {
          int64_t v = demand_pc[((size_t)q * PC + pc) * D + d];
          if (in->queue_has_limit && in->queue_has_limit[(size_t)q * PC + pc] && in->queue_limit) {
            const int64_t lim = in->queue_limit[((size_t)q * PC + pc) * D + d];
            if (lim < v) v = lim;
          }
          der_cdemand[(size_t)q * D + d] += v;
        }
  }

  // AddQueueSchedulingContext (scheduling.go:104-156) + UpdateFairShares (:174-182)
  void build_queue_contexts() {
    if (!in->queue_allocated_by_pc || !in->queue_constrained_demand) derive_snapshot();
    qctx.resize(Q);
    for (uint32_t q = 0; q < Q; ++q) {
      QueueCtx& c = qctx[q];
      c.weight = in->queue_weight[q];
      c.allocated.assign(D, 0);
      c.allocated_by_pc.assign((size_t)PC * D, 0);
      c.penalty.assign(D, 0);
      for (int pc = 0; pc < PC; ++pc)
        for (int d = 0; d < D; ++d) {
          int64_t v = in->queue_allocated_by_pc ? in->queue_allocated_by_pc[((size_t)q * PC + pc) * D + d] : der_alloc_pc[((size_t)q * PC + pc) * D + d];
          c.allocated_by_pc[(size_t)pc * D + d] = v;
          c.allocated[d] += v;
        }
      if (in->queue_short_job_penalty)
        for (int d = 0; d < D; ++d) c.penalty[d] = in->queue_short_job_penalty[(size_t)q * D + d];
      c.tokens = in->queue_limiter_tokens ? in->queue_limiter_tokens[q] : 0;
      c.burst = in->queue_limiter_burst ? in->queue_limiter_burst[q] : 0;
      c.inf = in->queue_limiter_is_inf ? in->queue_limiter_is_inf[q] != 0 : true;
      weight_sum += c.weight;  // NB: reference sums in Go map order; we use name order
    }
    update_fair_shares();
  }

  double unweighted_cost(const int64_t* a) const {
    return drf_unweighted(D, in->total_resources, in->drf_multipliers, a);
  }

  // updateFairShares, scheduling/context/scheduling.go:252-332 (queues already in name order)
  void update_fair_shares() {
    struct Info {
      double cds, weight, fair, capped = 0, uncapped = 0, spare = 0;
      bool achieved = false;
    };
    std::vector<Info> qi(Q);
    bool total_all_zero = true;
    for (int d = 0; d < D; ++d)
      if (in->total_resources[d] != 0) total_all_zero = false;
    for (uint32_t q = 0; q < Q; ++q) {
      double cds = 1.0;
      if (!total_all_zero) {
        std::vector<int64_t> cd(D, 0);
        for (int d = 0; d < D; ++d)
          cd[d] = in->queue_constrained_demand ? in->queue_constrained_demand[(size_t)q * D + d] : der_cdemand[(size_t)q * D + d];
        cds = unweighted_cost(cd.data());
      }
      qi[q].cds = cds;
      qi[q].weight = qctx[q].weight;
      qi[q].fair = qctx[q].weight / weight_sum;
    }
    double unallocated = 1.0;
    for (int i = 0; i < 10 && unallocated > 0.01; ++i) {
      double total_weight = 0.0;
      for (auto& q : qi)
        if (!q.achieved) total_weight += q.weight;
      for (auto& q : qi) {
        double tw = total_weight;
        if (q.achieved) tw += q.weight;
        q.uncapped += (q.weight / tw) * (unallocated - q.spare);
      }
      if (total_weight <= 0.0) break;
      for (auto& q : qi)
        if (!q.achieved) q.capped += (q.weight / total_weight) * unallocated;
      unallocated = 0.0;
      for (auto& q : qi) {
        double spare = q.capped - q.cds;
        if (spare > 0) {
          q.capped = q.cds;
          q.achieved = true;
          q.spare = spare;
          unallocated += spare;
        } else {
          q.spare = 0;
        }
      }
    }
    for (uint32_t q = 0; q < Q; ++q) {
      qctx[q].fair_share = qi[q].fair;
      qctx[q].demand_capped = qi[q].capped;
      qctx[q].uncapped = qi[q].uncapped;
    }
  }

  // SchedulingOrderCompare, jobdb/comparison.go:49-107
  bool order_less(uint32_t a, uint32_t b) const {
    if (a == b) return false;
    bool aa = in->job_node[a] != NONE, ba = in->job_node[b] != NONE;  // active run
    if (aa != ba) return aa;
    int32_t pa = db.pc_of(a).priority, pb = db.pc_of(b).priority;
    if (pa != pb) return pa > pb;
    if (in->job_queue_priority[a] != in->job_queue_priority[b])
      return in->job_queue_priority[a] < in->job_queue_priority[b];
    if (aa && ba && in->job_active_run_timestamp[a] != in->job_active_run_timestamp[b])
      return in->job_active_run_timestamp[a] < in->job_active_run_timestamp[b];
    if (in->job_submit_time[a] != in->job_submit_time[b]) return in->job_submit_time[a] < in->job_submit_time[b];
    return in->job_id_rank[a] < in->job_id_rank[b];
  }

  void build_queued_order() {
    queued_by_queue.resize(Q);
    if (in->queued_order && in->queued_start) {
      for (uint32_t q = 0; q < Q; ++q)
        for (uint32_t i = in->queued_start[q]; i < in->queued_start[q + 1]; ++i)
          queued_by_queue[q].push_back(in->queued_order[i]);
    } else {  // jobdb.Txn.QueuedJobs order, jobdb.go:878-896
      for (uint32_t j = 0; j < J; ++j)
        if (in->job_node[j] == NONE && in->job_queue[j] != NONE) queued_by_queue[in->job_queue[j]].push_back(j);
      for (auto& v : queued_by_queue) std::sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) { return order_less(a, b); });
    }
  }

  // ---- SchedulingContext accounting (context/scheduling.go:381-491, queue.go:231-331) ----
  void vadd(std::vector<int64_t>& v, const int64_t* r, int sign, size_t off = 0) {
    for (int d = 0; d < D; ++d) v[off + d] += sign * r[d];
  }
  // qctx.addJobSchedulingContext + sctx.AddJobSchedulingContext
  bool add_job(JobCtx& jc) {
    uint32_t j = jc.job, q = in->job_queue[j];
    if (q == NONE) fail(ARMADA_E_INTERNAL, "failed adding job to scheduling context: no context for queue");
    QueueCtx& c = qctx[q];
    if (successful[j]) fail(ARMADA_E_INTERNAL, "failed adding job to queue: job already marked successful");
    unsuccessful[j] = 0;
    bool ev = evicted_by_id[j];
    const int64_t* rl = req_of(j);
    if (jc.is_successful()) {
      vadd(c.allocated_by_pc, rl, +1, (size_t)pc_index(j) * D);
      vadd(c.allocated, rl, +1);
      if (ev) {
        evicted_by_id[j] = 0;
        rescheduled[j] = 1;
        vadd(evicted_resources, rl, -1);
        --num_evicted_jobs;
      } else {
        successful[j] = 1;
        vadd(scheduled_resources, rl, +1);
        ++num_scheduled_jobs;
      }
    } else {
      unsuccessful[j] = 1;
      unsuccessful_reason[j] = jc.reason;
    }
    return ev;
  }
  bool add_gang(std::vector<JobCtx*>& g) {  // AddGangSchedulingContext :381-397
    bool all_ev = true, all_succ = true;
    for (JobCtx* jc : g) {
      bool ev = add_job(*jc);
      all_ev = all_ev && ev;
      all_succ = all_succ && jc->is_successful();
    }
    if (all_succ && !all_ev) ++num_scheduled_gangs;
    return all_ev;
  }
  bool evict_job_ctx(uint32_t j) {  // sctx.EvictJob :470-491 + qctx.evictJob queue.go:296-331
    uint32_t q = in->job_queue[j];
    if (q == NONE) fail(ARMADA_E_INTERNAL, "failed evicting job: no context for queue");
    QueueCtx& c = qctx[q];
    if (unsuccessful[j]) fail(ARMADA_E_INTERNAL, "failed evicting job from queue: job already marked unsuccessful");
    if (evicted_by_id[j]) fail(ARMADA_E_INTERNAL, "failed evicting job from queue: job already marked evicted");
    const int64_t* rl = req_of(j);
    bool sched = successful[j], resched = rescheduled[j];
    if (sched || resched) {
      successful[j] = 0;
      rescheduled[j] = 0;
    } else {
      evicted_by_id[j] = 1;
    }
    vadd(c.allocated_by_pc, rl, -1, (size_t)pc_index(j) * D);
    vadd(c.allocated, rl, -1);
    if (sched) {
      vadd(scheduled_resources, rl, -1);
      --num_scheduled_jobs;
    } else {
      vadd(evicted_resources, rl, +1);
      ++num_evicted_jobs;
    }
    return sched;
  }
  void evict_gang_ctx(std::vector<JobCtx*>& g) {  // EvictGang :426-439
    bool all = true;
    for (JobCtx* jc : g) all = evict_job_ctx(jc->job) && all;
    if (all) --num_scheduled_gangs;
  }

  // ---- constraints (scheduling/constraints/constraints.go:113-178) ----------------------
  bool exceeds(const int64_t* a, const int64_t* b) const {  // ResourceList.Exceeds
    for (int d = 0; d < D; ++d)
      if (a[d] > b[d]) return true;
    return false;
  }
  uint8_t check_round_constraints() const {
    if (in->has_round_limit && exceeds(scheduled_resources.data(), in->max_resources_to_schedule))
      return ARMADA_REASON_MAX_RESOURCES_SCHEDULED;
    return ARMADA_REASON_NONE;
  }
  uint8_t check_job_constraints(uint32_t q, uint32_t pc, int cardinality) const {
    const QueueCtx& c = qctx[q];
    if (in->queue_cordoned && in->queue_cordoned[q]) return ARMADA_REASON_QUEUE_CORDONED;
    double tokens = global_tokens;
    if (tokens <= 0) return ARMADA_REASON_GLOBAL_RATE_LIMIT;
    if (in->global_limiter_burst < cardinality) return ARMADA_REASON_GANG_EXCEEDS_GLOBAL_BURST;
    if (tokens < (double)cardinality) return ARMADA_REASON_GLOBAL_RATE_LIMIT_GANG;
    tokens = c.tokens;
    if (tokens <= 0) return ARMADA_REASON_QUEUE_RATE_LIMIT;
    if (c.burst < cardinality) return ARMADA_REASON_GANG_EXCEEDS_QUEUE_BURST;
    if (tokens < (double)cardinality) return ARMADA_REASON_QUEUE_RATE_LIMIT_GANG;
    // soft wall-clock limits (:160-169) are nondeterministic and not modelled (disabled = 0)
    if (in->queue_has_limit && in->queue_has_limit[(size_t)q * PC + pc]) {
      const int64_t* lim = in->queue_limit + ((size_t)q * PC + pc) * D;
      if (exceeds(c.allocated_by_pc.data() + (size_t)pc * D, lim)) return ARMADA_REASON_MAX_RESOURCES_PER_QUEUE;
    }
    return ARMADA_REASON_NONE;
  }
  static bool is_terminal(uint8_t r) {  // IsTerminalUnschedulableReason :63-67
    return r == ARMADA_REASON_MAX_RESOURCES_SCHEDULED || r == ARMADA_REASON_GLOBAL_RATE_LIMIT;
  }
  static bool is_queue_terminal(uint8_t r) {  // IsTerminalQueueUnschedulableReason :71-75
    return r == ARMADA_REASON_QUEUE_RATE_LIMIT || r == ARMADA_REASON_QUEUE_CORDONED;
  }
  static bool is_property_of_gang(uint8_t r) {  // UnschedulableReasonIsPropertyOfGang :57-59
    return r == ARMADA_REASON_GANG_EXCEEDS_GLOBAL_BURST || r == ARMADA_REASON_JOB_DOES_NOT_FIT ||
           r == ARMADA_REASON_GANG_DOES_NOT_FIT;
  }

  // ---- GangScheduler (scheduling/gang_scheduler.go:100-254) -----------------------------
  struct Gang {  // GangSchedulingContext, context/gang.go:9-49
    std::vector<JobCtx*> jctxs;
    std::vector<int64_t> total;
    bool all_evicted = true;
    uint32_t queue = NONE;
    uint32_t pc = 0;
  };
  Gang make_gang(std::vector<JobCtx*> jctxs) {
    Gang g;
    g.total.assign(D, 0);
    for (JobCtx* jc : jctxs) {
      g.all_evicted = g.all_evicted && jc->is_evicted;
      const int64_t* r = req_of(jc->job);
      for (int d = 0; d < D; ++d) g.total[d] += r[d];
    }
    g.queue = in->job_queue[jctxs[0]->job];
    g.pc = pc_index(jctxs[0]->job);
    g.jctxs = std::move(jctxs);
    return g;
  }

  // sctx.IsWithinFloatingResourceLimits (context/scheduling.go:493-517) + FloatingResourceTypes.WithinLimits
  // (floatingresources/floating_resource_types.go:60-72).  sctx.Allocated = the sum of the queues' allocations,
  // this gang included (AddGangSchedulingContext ran).  Jobs of other pools ("away" in the cross-pool sense) are
  // not expressible in the ABI.
  uint8_t check_floating_resources(const Gang& g) const {
    if (!in->floating_resource_mask) return ARMADA_REASON_NONE;
    bool requests = false;
    for (JobCtx* jc : g.jctxs) {
      const int64_t* r = req_of(jc->job);
      for (int d = 0; d < D; ++d)
        if (((in->floating_resource_mask >> d) & 1u) && r[d] > 0) requests = true;
    }
    if (!requests) return ARMADA_REASON_NONE;
    if (!in->floating_limits_configured) return ARMADA_REASON_FLOATING_RESOURCES;
    for (int d = 0; d < D; ++d) {
      if (!((in->floating_resource_mask >> d) & 1u)) continue;
      int64_t a = 0;
      for (uint32_t q = 0; q < Q; ++q) a += qctx[q].allocated[d];
      if (a > in->floating_limit[d]) return ARMADA_REASON_FLOATING_RESOURCES;
    }
    return ARMADA_REASON_NONE;
  }
  // tryScheduleGang (gang_scheduler.go:225-238)
  bool try_schedule_gang(Gang& g, uint8_t* reason) {
    db.begin();
    bool ok = db.schedule_many(g.jctxs);
    if (ok) {
      db.commit();
    } else {
      db.abort();
      *reason = g.jctxs.size() > 1 ? ARMADA_REASON_GANG_DOES_NOT_FIT : ARMADA_REASON_JOB_DOES_NOT_FIT;
    }
    return ok;
  }
  void set_uniformity_slot(Gang& g, int64_t slot) {  // addNodeSelectorToGctx (:256-260)
    db.uniformity_slot = slot;
    for (JobCtx* jc : g.jctxs) {
      jc->row = db.base_row(in->job_class[jc->job]);
      if (slot >= 0) jc->has_additional = true;
    }
  }
  // trySchedule (gang_scheduler.go:150-223): one attempt per value of the gang's node-uniformity label, each in a
  // transaction that is rolled back unless it is the best possible fit; the best value is then scheduled for real.
  // Values are tried in slot (string) order — the reference ranges over a Go map.
  bool try_schedule(Gang& g, uint8_t* reason) {
    const uint32_t gi = in->job_gang[g.jctxs[0]->job];
    const uint32_t label = (g.jctxs.size() > 1 && gi != NONE && in->gang_uniformity_label) ? in->gang_uniformity_label[gi] : NONE;
    if (label == NONE) return try_schedule_gang(g, reason);
    if (label == ARMADA_LABEL_NOT_INDEXED) {
      *reason = ARMADA_REASON_UNIFORMITY_LABEL_NOT_INDEXED;
      return false;
    }
    const uint32_t v0 = in->uniformity_value_start[label], v1 = in->uniformity_value_start[label + 1];
    if (v0 == v1) {
      *reason = ARMADA_REASON_NO_NODES_WITH_UNIFORMITY_LABEL;
      return false;
    }
    int64_t best = -1;
    double best_mean = 0;  // GangSchedulingFit (context/gang.go:82-110): every attempt that succeeds places the whole gang
    for (uint32_t v = v0; v < v1; ++v) {
      set_uniformity_slot(g, v);
      db.begin();
      const bool ok = db.schedule_many(g.jctxs);
      if (ok) {
        int32_t total = 0;
        for (JobCtx* jc : g.jctxs) total += jc->p_preempted_at;
        const double mean = (double)total / (double)g.jctxs.size();
        if (mean == (double)MIN_PRIORITY || ((best < 0 || best_mean > mean) && v + 1 == v1)) {
          db.commit();  // best possible, or the best so far with nothing left to try
          db.uniformity_slot = -1;
          return true;
        }
        if (best < 0 || best_mean > mean) {
          best = v;
          best_mean = mean;
        }
      }
      db.abort();
    }
    if (best < 0) {
      set_uniformity_slot(g, -1);
      *reason = ARMADA_REASON_GANG_FITS_NO_UNIFORMITY_VALUE;
      return false;
    }
    set_uniformity_slot(g, best);
    const bool ok = try_schedule_gang(g, reason);
    db.uniformity_slot = -1;
    return ok;
  }

  bool gang_schedule(Gang& g, bool skip_key_check, uint8_t* reason_out) {
    uint8_t reason = ARMADA_REASON_NONE;
    bool ok = true;
    if (!g.all_evicted) {
      reason = check_round_constraints();
      if (reason != ARMADA_REASON_NONE) {
        // gang_scheduler.go:102-106: returns BEFORE the deferred bookkeeping func is
        // registered (:109), so nothing is recorded for this gang (jobs stay untouched).
        *reason_out = reason;
        return false;
      }
    }
    add_gang(g.jctxs);
    bool added = true;
    if (!g.all_evicted) {
      reason = check_job_constraints(g.queue, g.pc, (int)g.jctxs.size());
      if (reason != ARMADA_REASON_NONE) ok = false;
    }
    if (ok) {
      reason = check_floating_resources(g);  // IsWithinFloatingResourceLimits (:143)
      if (reason != ARMADA_REASON_NONE) ok = false;
    }
    if (ok) ok = try_schedule(g, &reason);
    if (ok) {
      if (!g.all_evicted) {  // Limiter.ReserveN (:118-123)
        if (!in->global_limiter_is_inf) global_tokens -= (double)g.jctxs.size();
        QueueCtx& c = qctx[g.queue];
        if (!c.inf) c.tokens -= (double)g.jctxs.size();
      }
      // updateGangSchedulingContextOnSuccess :43-61
      for (JobCtx* jc : g.jctxs)
        if (!jc->is_successful()) evict_job_ctx(jc->job);
      for (JobCtx* jc : g.jctxs) {
        last_jctx[jc->job] = *jc;
        ++stats.placements;
      }
    } else {
      gang_failure(g, added, reason, skip_key_check);
    }
    *reason_out = reason;
    return ok;
  }
  // updateGangSchedulingContextOnFailure :63-98
  void gang_failure(Gang& g, bool added, uint8_t reason, bool skip_key_check) {
    if (added) evict_gang_ctx(g.jctxs);
    for (JobCtx* jc : g.jctxs) {  // jctx.Fail
      jc->reason = reason;
      if (jc->has_pctx) {
        jc->p_node = NONE;
        jc->p_method = ARMADA_METHOD_NONE;
      }
    }
    add_gang(g.jctxs);
    for (JobCtx* jc : g.jctxs) last_jctx[jc->job] = *jc;
    if (!skip_key_check && g.jctxs.size() == 1 && is_property_of_gang(reason)) {
      JobCtx* jc = g.jctxs[0];
      uint32_t c = in->job_class[jc->job];
      if (!jc->has_additional && in->class_key_valid[c] && !unfeasible_key[c]) {
        unfeasible_key[c] = reason;
        ++num_unfeasible;
      }
    }
  }

  // ---- job / gang iterators (scheduling/jobiteration.go, queue_scheduler.go:277-383) ------
  struct GangIt {  // QueuedGangIterator over MultiJobsIterator(evicted, queued)
    uint32_t queue = NONE;
    std::vector<JobCtx*> evicted;  // InMemoryJobIterator
    size_t ei = 0;
    const std::vector<uint32_t>* queued = nullptr;  // QueuedJobsIterator (nullptr = none)
    size_t qi = 0;
    bool queued_only_evicted = false;
    std::map<uint32_t, std::vector<JobCtx*>> by_gang;
    uint32_t max_lookback = 0;
    bool skip_known_unschedulable = false;
    uint32_t jobs_seen = 0;
    bool only_yield_evicted = false;
    bool has_next = false;
    Gang next;
  };
  std::vector<std::unique_ptr<JobCtx>> jctx_pool;
  JobCtx* new_jctx(uint32_t job) {  // JobSchedulingContextFromJob, context/job.go:149-158
    jctx_pool.emplace_back(new JobCtx());
    JobCtx* jc = jctx_pool.back().get();
    jc->job = job;
    jc->row = in->class_static_row[in->job_class[job]];
    uint32_t g = in->job_gang[job];
    jc->current_gang_cardinality = g == NONE ? 1 : (int)in->gang_cardinality[g];
    return jc;
  }
  JobCtx* gang_it_next_job(GangIt& it) {  // MultiJobsIterator.Next, jobiteration.go:166-180
    if (it.ei < it.evicted.size()) return it.evicted[it.ei++];
    if (it.queued && !it.queued_only_evicted && it.qi < it.queued->size()) return new_jctx((*it.queued)[it.qi++]);
    return nullptr;
  }
  void gang_it_only_yield_evicted(GangIt& it) {  // queue_scheduler.go:303-309
    it.only_yield_evicted = true;
    it.queued_only_evicted = true;
    if (it.has_next && !it.next.all_evicted) it.has_next = false;
  }
  bool gang_it_peek(GangIt& it) {  // Peek :315-371
    if (it.has_next) return true;
    for (;;) {
      if (it.max_lookback != 0 && !it.only_yield_evicted && it.jobs_seen >= it.max_lookback)
        gang_it_only_yield_evicted(it);  // stopYieldingNewJobsIfLimitHit :373-383
      JobCtx* jc = gang_it_next_job(it);
      if (!jc) return false;
      if (!jc->is_evicted) ++it.jobs_seen;
      if (it.skip_known_unschedulable && num_unfeasible > 0) {
        uint32_t c = in->job_class[jc->job];
        if (!jc->has_additional && in->class_key_valid[c] && unfeasible_key[c]) {
          jc->reason = unfeasible_key[c];
          jc->has_pctx = true;  // shares the failed jctx's pctx (NodeId == "")
          jc->p_node = NONE;
          jc->p_method = ARMADA_METHOD_NONE;
          add_job(*jc);
          last_jctx[jc->job] = *jc;
          continue;
        }
      }
      uint32_t g = in->job_gang[jc->job];
      if (g != NONE) {
        auto& v = it.by_gang[g];
        v.push_back(jc);
        if ((int)v.size() == jc->current_gang_cardinality) {
          std::vector<JobCtx*> members = std::move(v);
          it.by_gang.erase(g);
          it.next = make_gang(std::move(members));
          it.has_next = true;
          return true;
        }
      } else {
        it.next = make_gang({jc});
        it.has_next = true;
        return true;
      }
    }
  }

  // ---- CostBasedCandidateGangIterator (queue_scheduler.go:388-593) -----------------------
  struct PQItem {
    uint32_t queue;
    GangIt* it;
    bool has_gang = false;
    double proposed = 0, current = 0, budget = 0, item_size = 0;
    int32_t pc_priority = 0;
  };
  struct CandidateIt {
    std::vector<PQItem> items;  // the heap contents
    bool consider_priority = false, prioritise_larger = false;
    bool only_yield_evicted = false;
    std::vector<uint8_t> only_yield_evicted_by_queue;
  };
  // QueueCandidateGangIteratorPQ.Less :628-674
  static bool pq_less(const CandidateIt& ci, const PQItem& a, const PQItem& b) {
    if (ci.consider_priority && a.pc_priority != b.pc_priority) return a.pc_priority > b.pc_priority;
    if (ci.prioritise_larger) {
      bool au = a.proposed <= a.budget, bu = b.proposed <= b.budget;
      if (au && bu) {
        if (a.current == b.current && a.item_size != b.item_size) return a.item_size > b.item_size;
        if (a.current != b.current) return a.current < b.current;
      } else if (!au && !bu) {
        if (a.proposed != b.proposed) return a.proposed < b.proposed;
      } else if (au) {
        return true;
      } else if (bu) {
        return false;
      }
    } else {
      if (a.proposed != b.proposed) return a.proposed < b.proposed;
    }
    return a.queue < b.queue;
  }
  void pq_update(PQItem& item) {  // updatePQItem :536-580
    item.has_gang = false;
    item.proposed = item.current = item.item_size = 0;
    if (!gang_it_peek(*item.it)) return;
    item.has_gang = true;
    Gang& g = item.it->next;
    const QueueCtx& c = qctx[g.queue];
    std::vector<int64_t> base(D), with(D);
    for (int d = 0; d < D; ++d) {
      base[d] = c.allocated[d] + c.penalty[d];
      with[d] = base[d] + g.total[d];
    }
    item.proposed = unweighted_cost(with.data()) / c.weight;
    item.current = unweighted_cost(base.data()) / c.weight;
    item.item_size = unweighted_cost(g.total.data()) * c.weight;
    int32_t pr = std::numeric_limits<int32_t>::max();
    for (JobCtx* jc : g.jctxs) {
      int32_t np = db.pc_of(jc->job).priority;
      if (jc->has_pctx) {
        np = jc->p_sched_at;
      } else if (in->job_node[jc->job] != NONE && in->job_scheduled_at_priority[jc->job] != ARMADA_NO_PRIORITY) {
        np = in->job_scheduled_at_priority[jc->job];
      }
      if (np < pr) pr = np;
    }
    item.pc_priority = pr;
  }
  int pq_top(const CandidateIt& ci) const {
    int best = -1;
    for (size_t i = 0; i < ci.items.size(); ++i)
      if (best < 0 || pq_less(ci, ci.items[i], ci.items[(size_t)best])) best = (int)i;
    return best;
  }
  void candidate_init(CandidateIt& ci, std::vector<GangIt>& its, bool consider_priority, bool prioritise_larger) {
    ci.consider_priority = consider_priority;
    ci.prioritise_larger = prioritise_larger;
    ci.only_yield_evicted_by_queue.assign(Q, 0);
    for (auto& it : its) {
      PQItem item;
      item.queue = it.queue;
      item.it = &it;
      item.budget = qctx[it.queue].demand_capped / qctx[it.queue].weight;  // :438
      pq_update(item);
      if (item.has_gang) ci.items.push_back(item);
    }
  }
  void candidate_clear(CandidateIt& ci) {  // Clear :495-506
    int top = pq_top(ci);
    if (top < 0) return;
    PQItem item = ci.items[(size_t)top];
    ci.items.erase(ci.items.begin() + top);
    item.it->has_next = false;
    pq_update(item);
    if (item.has_gang) ci.items.push_back(item);
  }
  void candidate_only_yield_evicted(CandidateIt& ci) {  // :446-469
    if (!ci.only_yield_evicted) {
      std::vector<PQItem> np;
      for (auto& item : ci.items) {
        gang_it_only_yield_evicted(*item.it);
        pq_update(item);
        if (item.has_gang) np.push_back(item);
      }
      ci.items = std::move(np);
    }
    ci.only_yield_evicted = true;
  }
  void candidate_only_yield_evicted_for_queue(CandidateIt& ci, uint32_t queue) {  // :471-491
    if (!ci.only_yield_evicted && !ci.only_yield_evicted_by_queue[queue]) {
      for (size_t i = 0; i < ci.items.size(); ++i) {
        if (ci.items[i].queue == queue) {
          gang_it_only_yield_evicted(*ci.items[i].it);
          pq_update(ci.items[i]);
          if (!ci.items[i].has_gang) ci.items.erase(ci.items.begin() + (long)i);
          break;
        }
      }
    }
    ci.only_yield_evicted_by_queue[queue] = 1;
  }

  // ---- QueueScheduler.Schedule (queue_scheduler.go:91-272) -------------------------------
  // Returns the jobs of successfully scheduled gangs.
  std::vector<uint32_t> queue_schedule(std::vector<GangIt>& its, bool skip_key_check, bool consider_priority,
                                       uint32_t* termination) {
    std::vector<uint32_t> scheduled;
    CandidateIt ci;
    candidate_init(ci, its, consider_priority, in->prefer_large_job_ordering != 0);
    uint32_t term = ARMADA_REASON_NONE;
    for (;;) {
      int top = pq_top(ci);
      if (top < 0) break;
      Gang g = ci.items[(size_t)top].it->next;  // Peek
      ++stats.loop_iterations;
      for (JobCtx* jc : g.jctxs) job_seq[jc->job] = (uint32_t)stats.loop_iterations;
      uint8_t reason = ARMADA_REASON_NONE;
      bool ok = gang_schedule(g, skip_key_check, &reason);
      if (in_first_pass)
        for (JobCtx* jc : g.jctxs) {
          job_seq_first[jc->job] = (uint32_t)stats.loop_iterations;
          job_reason_first[jc->job] = ok ? (uint8_t)ARMADA_REASON_NONE : reason;
        }
      candidate_clear(ci);
      if (ok) {
        for (JobCtx* jc : g.jctxs)
          if (jc->pctx_successful()) scheduled.push_back(jc->job);
      } else if (is_terminal(reason)) {
        term = reason;
        candidate_only_yield_evicted(ci);
      } else if (is_queue_terminal(reason)) {
        candidate_only_yield_evicted_for_queue(ci, g.queue);
      }
    }
    if (term == ARMADA_REASON_NONE) term = ARMADA_REASON_NO_REMAINING_CANDIDATES;
    *termination = term;
    return scheduled;
  }

  // ---- Evictor (scheduling/eviction.go:186-273) ------------------------------------------
  struct EvictResult {
    std::vector<JobCtx*> evicted;  // EvictedJctxsByJobId
  };
  template <class NodeFilter, class JobFilter>
  void evictor_run(EvictResult& res, NodeFilter node_filter, JobFilter job_filter) {
    for (uint32_t n : db.nodes_by_id) {  // NodesIterator: "nodes" table in id order
      if (!node_filter(n)) continue;
      std::vector<uint32_t> jobs;
      for (uint32_t j : db.node_jobs[n])
        if (!db.evicted_on_node[j] && job_filter(j)) jobs.push_back(j);
      for (uint32_t j : jobs) db.evict(j);  // EvictJobsFromNode, nodedb.go:960-971
      for (uint32_t j : jobs) {
        JobCtx* jc = new_jctx(j);
        jc->is_evicted = true;
        jc->assigned_node = n;      // SetAssignedNode (+ nodeIdLabel selector)
        jc->has_additional = true;  // + tolerations for the node's taints
        res.evicted.push_back(jc);
      }
    }
  }

  // PreemptingQueueScheduler.evict, preempting_queue_scheduler.go:287-349
  template <class NodeFilter, class JobFilter>
  EvictResult evict(NodeFilter nf, JobFilter jf, std::vector<std::vector<JobCtx*>>& repo) {
    EvictResult res;
    evictor_run(res, nf, jf);
    // evictGangs :353-420: evict the remaining members of partially evicted gangs
    {
      std::set<uint32_t> gang_jobs, gang_nodes, seen;
      for (JobCtx* jc : res.evicted) {
        uint32_t g = in->job_gang[jc->job];
        if (g == NONE || seen.count(g)) continue;
        const auto& active = gang_members[g];  // getActiveGangJobs
        for (uint32_t gj : active) {
          gang_jobs.insert(gj);
          if (in->job_node[gj] == NONE) {  // queued: node from this round's scheduling context
            if (successful[gj] && last_jctx[gj].pctx_successful()) gang_nodes.insert(last_jctx[gj].p_node);
            continue;
          }
          gang_nodes.insert(in->job_node[gj]);
        }
        seen.insert(g);
      }
      if (!gang_jobs.empty() && !gang_nodes.empty()) {  // NewFilteredEvictor
        EvictResult gres;
        evictor_run(gres, [&](uint32_t n) { return gang_nodes.count(n) > 0; },
                    [&](uint32_t j) { return gang_jobs.count(j) > 0; });
        for (JobCtx* jc : gres.evicted) res.evicted.push_back(jc);
      }
    }
    // setEvictedGangCardinality :458-479
    for (JobCtx* jc : res.evicted) {
      uint32_t g = in->job_gang[jc->job];
      if (g != NONE) jc->current_gang_cardinality = (int)gang_members[g].size();
    }
    for (JobCtx* jc : res.evicted) evict_job_ctx(jc->job);
    // InMemoryJobRepository.EnqueueMany + sortQueue, jobiteration.go:72-95
    repo.assign(Q, {});
    for (JobCtx* jc : res.evicted) repo[in->job_queue[jc->job]].push_back(jc);
    for (auto& v : repo)
      std::sort(v.begin(), v.end(), [&](JobCtx* a, JobCtx* b) { return order_less(a->job, b->job); });
    // nodeDb.Reset :260-274
    db.ev_clear();
    std::fill(db.evicted_index_of_job.begin(), db.evicted_index_of_job.end(), -1);
    add_evicted_jobs_to_nodedb(repo);
    return res;
  }

  // addEvictedJobsToNodeDb :584-633 — drain a cost-ordered iterator with FROZEN allocations
  void add_evicted_jobs_to_nodedb(std::vector<std::vector<JobCtx*>>& repo) {
    std::vector<GangIt> its(Q);
    for (uint32_t q = 0; q < Q; ++q) {
      its[q].queue = q;
      its[q].evicted = repo[q];
      its[q].max_lookback = 0;
      its[q].skip_known_unschedulable = false;
    }
    CandidateIt ci;
    candidate_init(ci, its, false, in->prefer_large_job_ordering != 0);
    int i = 0;
    for (;;) {
      int top = pq_top(ci);
      if (top < 0) break;
      Gang& g = ci.items[(size_t)top].it->next;
      for (JobCtx* jc : g.jctxs) {
        if (db.evicted_index_of_job[jc->job] >= 0) fail(ARMADA_E_INTERNAL, "tried to insert evicted job with duplicate index");
        db.ev_insert(i, jc->job);
        db.evicted_index_of_job[jc->job] = i;
        ++i;
      }
      candidate_clear(ci);
    }
  }

  // PreemptingQueueScheduler.schedule :704-763
  std::vector<uint32_t> schedule_pass(std::vector<std::vector<JobCtx*>>& repo, bool with_queued, bool skip_key_check,
                                      bool consider_priority, uint32_t* termination) {
    std::vector<GangIt> its(Q);
    for (uint32_t q = 0; q < Q; ++q) {
      its[q].queue = q;
      its[q].evicted = repo[q];
      its[q].queued = with_queued ? &queued_by_queue[q] : nullptr;
      its[q].max_lookback = in->max_queue_lookback;
      its[q].skip_known_unschedulable = true;
    }
    std::fill(unfeasible_key.begin(), unfeasible_key.end(), 0);  // ClearUnfeasibleSchedulingKeys :731
    num_unfeasible = 0;
    return queue_schedule(its, skip_key_check, consider_priority, termination);
  }

  // PreemptingQueueScheduler.Schedule :84-285
  void run(ArmadaRoundOutput* out) {
    db.bind_running_jobs();
    std::vector<uint8_t> preempted(J, 0), scheduled(J, 0), sched_and_evicted(J, 0);
    std::vector<uint32_t> assigned_node(J, NONE);  // jctx.AssignedNode of the evicted jctx

    // 1. evict for resource balancing (:94-136)
    std::vector<uint8_t> evict_queue(Q, 0);
    for (uint32_t q = 0; q < Q; ++q) {
      const QueueCtx& c = qctx[q];
      double actual = unweighted_cost(c.allocated.data());
      double fair = std::max(c.demand_capped, c.fair_share);  // math.Max
      if (std::isnan(c.demand_capped) || std::isnan(c.fair_share)) fair = std::nan("");
      if (in->protect_uncapped_adjusted_fair_share) fair = c.uncapped;
      double fraction = actual / fair;
      evict_queue[q] = !(fraction <= in->protected_fraction_of_fair_share);
    }
    std::vector<std::vector<JobCtx*>> repo;
    EvictResult ev1 = evict([&](uint32_t n) { return !db.node_jobs[n].empty(); },
                            [&](uint32_t j) {
                              uint32_t q = in->job_queue[j];
                              if (q == NONE) return false;            // invalid_queue
                              if (!db.pc_of(j).preemptible) return false;  // job_not_preemptible
                              return evict_queue[q] != 0;
                            },
                            repo);
    stats.evicted_pass1 = ev1.evicted.size();
    for (JobCtx* jc : ev1.evicted) {
      preempted[jc->job] = 1;
      assigned_node[jc->job] = jc->assigned_node;
    }

    // 2. re-schedule evicted + schedule new (:140-164)
    uint32_t term1 = 0;
    in_first_pass = true;
    std::vector<uint32_t> res1 = schedule_pass(repo, true, false, false, &term1);
    in_first_pass = false;
    termination_reason = term1;
    for (uint32_t j : res1) {
      if (preempted[j]) preempted[j] = 0;
      else scheduled[j] = 1;
    }

    // 3. evict jobs on oversubscribed nodes (:166-194, eviction.go:132-184)
    std::vector<uint8_t> oversub;
    EvictResult ev2 = evict(
        [&](uint32_t n) {
          oversub.assign((size_t)db.PL, 0);
          bool any = false;
          for (int p = 0; p < db.PL; ++p) {
            if (in->priorities[p] < 0) continue;
            for (int d = 0; d < D; ++d)
              if (db.A(p, d, n) < 0) {
                oversub[(size_t)p] = 1;
                any = true;
              }
          }
          return any;
        },
        [&](uint32_t j) {
          if (in->job_queue[j] == NONE) return false;
          if (!db.pc_of(j).preemptible) return false;
          if (!db.has_sched_prio[j]) return false;
          for (int p = 0; p < db.PL; ++p)
            if (in->priorities[p] == db.sched_prio[j]) return oversub[(size_t)p] != 0;
          return false;
        },
        repo);
    stats.evicted_pass2 = ev2.evicted.size();
    for (JobCtx* jc : ev2.evicted) {
      uint32_t j = jc->job;
      assigned_node[j] = jc->assigned_node;
      if (scheduled[j]) {
        scheduled[j] = 0;
        sched_and_evicted[j] = 1;
      } else {
        preempted[j] = 1;
      }
    }

    // 4. second schedule pass, evicted jobs only (:196-219)
    if (!ev2.evicted.empty()) {
      uint32_t term2 = 0;
      std::vector<uint32_t> res2 = schedule_pass(repo, false, true, true, &term2);
      for (uint32_t j : res2) {
        if (preempted[j]) preempted[j] = 0;
        else scheduled[j] = 1;
        sched_and_evicted[j] = 0;
      }
    }

    // 6. unbindJobs(preempted ∪ scheduledAndEvicted) :256, :766-789
    for (uint32_t j = 0; j < J; ++j)
      if (preempted[j] || sched_and_evicted[j]) db.unbind(j, assigned_node[j]);

    stats.probes = db.stat_probes;
    stats.fair_preemption_scans = db.stat_fair_scans;
    write_output(out, preempted, scheduled, sched_and_evicted, assigned_node);
  }

  void write_output(ArmadaRoundOutput* out, const std::vector<uint8_t>& preempted, const std::vector<uint8_t>& scheduled,
                    const std::vector<uint8_t>& sae, const std::vector<uint32_t>& assigned_node) {
    if (!out) return;
    uint32_t ns = 0, np = 0;
    for (uint32_t j = 0; j < J; ++j) {
      uint8_t st = ARMADA_JOB_NONE;
      uint32_t node = NONE;
      const JobCtx& jc = last_jctx[j];
      if (preempted[j]) {
        st = ARMADA_JOB_PREEMPTED;
        node = assigned_node[j];
        ++np;
      } else if (scheduled[j]) {
        st = ARMADA_JOB_SCHEDULED;
        node = jc.p_node;
        ++ns;
      } else if (sae[j]) {
        st = ARMADA_JOB_SCHEDULED_AND_EVICTED;
        node = assigned_node[j];
      } else if (rescheduled[j]) {
        st = ARMADA_JOB_RESCHEDULED;
        node = jc.p_node;
      } else if (unsuccessful[j]) {
        st = ARMADA_JOB_FAILED;
      }
      if (out->job_state) out->job_state[j] = st;
      if (out->job_node) out->job_node[j] = node;
      // Only placement-carrying states expose pod-scheduling details (keeps the parity surface
      // to what SchedulingResult consumers read, scheduling_algo.go:808-823).
      bool has = jc.job != NONE && jc.has_pctx && st != ARMADA_JOB_NONE && st != ARMADA_JOB_FAILED;
      if (out->job_scheduled_at_priority) out->job_scheduled_at_priority[j] = has ? jc.p_sched_at : ARMADA_NO_PRIORITY;
      if (out->job_preempted_at_priority) out->job_preempted_at_priority[j] = has ? jc.p_preempted_at : ARMADA_NO_PRIORITY;
      if (out->job_method) out->job_method[j] = has ? jc.p_method : (uint8_t)ARMADA_METHOD_NONE;
      if (out->job_reason) out->job_reason[j] = unsuccessful[j] ? unsuccessful_reason[j] : (uint8_t)ARMADA_REASON_NONE;
      if (out->job_seq) out->job_seq[j] = job_seq[j];
      if (out->job_seq_first_pass) out->job_seq_first_pass[j] = job_seq_first[j];
      if (out->job_reason_first_pass) out->job_reason_first_pass[j] = job_reason_first[j];
      if (out->job_excluded_nodes && in->collect_excluded_nodes) {
        const bool single_failed = st == ARMADA_JOB_FAILED && jc.job != NONE && jc.has_pctx && in->job_gang[j] == NONE;
        for (uint32_t k = 0; k < ARMADA_EXCLUDED_KINDS; ++k) out->job_excluded_nodes[(size_t)j * ARMADA_EXCLUDED_KINDS + k] = single_failed ? jc.excl[k] : 0u;
      }
    }
    if (out->node_alloc) std::memcpy(out->node_alloc, db.alloc.data(), db.alloc.size() * sizeof(int64_t));
    for (uint32_t q = 0; q < Q; ++q) {
      if (out->queue_allocated)
        for (int d = 0; d < D; ++d) out->queue_allocated[(size_t)q * D + d] = qctx[q].allocated[d];
      if (out->queue_allocated_by_pc)
        for (size_t i = 0; i < (size_t)PC * D; ++i) out->queue_allocated_by_pc[(size_t)q * PC * D + i] = qctx[q].allocated_by_pc[i];
      if (out->queue_fair_share) {
        out->queue_fair_share[(size_t)q * 3 + 0] = qctx[q].fair_share;
        out->queue_fair_share[(size_t)q * 3 + 1] = qctx[q].demand_capped;
        out->queue_fair_share[(size_t)q * 3 + 2] = qctx[q].uncapped;
      }
    }
    for (int d = 0; d < D; ++d) {
      if (out->scheduled_resources) out->scheduled_resources[d] = scheduled_resources[d];
      if (out->evicted_resources) out->evicted_resources[d] = evicted_resources[d];
    }
    out->num_scheduled_jobs = (uint32_t)num_scheduled_jobs;
    out->num_scheduled_gangs = (uint32_t)num_scheduled_gangs;
    out->num_evicted_jobs = (int32_t)num_evicted_jobs;
    out->termination_reason = termination_reason;
    out->num_result_scheduled = ns;
    out->num_result_preempted = np;
  }
};

// ---- input validation shared by both entry points -------------------------------------------
void validate(const ArmadaRoundInput* in) {
  if (!in) fail(ARMADA_E_INVALID, "null input");
  if (in->abi_version != ARMADA_ABI_VERSION) fail(ARMADA_E_INVALID, "abi version mismatch");
  if (in->num_resources == 0 || in->num_resources > ARMADA_MAX_RESOURCES) fail(ARMADA_E_INVALID, "num_resources");
  if (in->num_indexed == 0 || in->num_indexed > in->num_resources) fail(ARMADA_E_INVALID, "num_indexed");
  for (uint32_t i = 0; i < in->num_indexed; ++i) {
    if (in->indexed_resource[i] >= in->num_resources) fail(ARMADA_E_INVALID, "indexed_resource");
    if (in->indexed_resolution[i] <= 0) fail(ARMADA_E_INVALID, "indexed_resolution must be > 0");
  }
  if (in->num_priorities < 2 || in->num_priorities > ARMADA_MAX_PRIORITIES) fail(ARMADA_E_INVALID, "num_priorities");
  if (in->priorities[0] != -1) fail(ARMADA_E_INVALID, "priorities[0] must be -1");
  for (uint32_t p = 1; p < in->num_priorities; ++p)
    if (in->priorities[p] <= in->priorities[p - 1]) fail(ARMADA_E_INVALID, "priorities not ascending");
  if (in->num_priority_classes == 0 || in->num_priority_classes > ARMADA_MAX_PRIORITY_CLASSES)
    fail(ARMADA_E_INVALID, "num_priority_classes");
  if (in->num_nodes && (!in->node_index || !in->node_id_rank || !in->node_type || !in->node_static_class ||
                        !in->node_flags || !in->node_total || !in->node_allocatable))
    fail(ARMADA_E_INVALID, "null node array");
  if (in->num_classes && (!in->class_request || !in->class_pc || !in->class_static_row || !in->class_away_row ||
                          !in->class_key_valid || !in->static_match || !in->type_match))
    fail(ARMADA_E_INVALID, "null class array");
  if (in->num_jobs && (!in->job_class || !in->job_queue || !in->job_queue_priority || !in->job_submit_time ||
                       !in->job_id_rank || !in->job_gang || !in->job_node || !in->job_scheduled_at_priority ||
                       !in->job_active_run_timestamp))
    fail(ARMADA_E_INVALID, "null job array");
  if (in->num_queues && !in->queue_weight) fail(ARMADA_E_INVALID, "null queue array");
  for (uint32_t n = 0; n < in->num_nodes; ++n) {
    if (in->node_type[n] >= in->num_node_types) fail(ARMADA_E_INVALID, "node_type out of range");
    if (in->node_static_class[n] >= in->num_static_classes) fail(ARMADA_E_INVALID, "node_static_class out of range");
  }
  for (uint32_t c = 0; c < in->num_classes; ++c) {
    if (in->class_pc[c] >= in->num_priority_classes) fail(ARMADA_E_INVALID, "class_pc out of range");
    if (in->class_static_row[c] >= in->num_static_rows) fail(ARMADA_E_INVALID, "class_static_row out of range");
  }
  for (uint32_t j = 0; j < in->num_jobs; ++j) {
    if (in->job_class[j] >= in->num_classes) fail(ARMADA_E_INVALID, "job_class out of range");
    if (in->job_queue[j] != NONE && in->job_queue[j] >= in->num_queues) fail(ARMADA_E_INVALID, "job_queue out of range");
    if (in->job_node[j] != NONE && in->job_node[j] >= in->num_nodes) fail(ARMADA_E_INVALID, "job_node out of range");
    if (in->job_gang[j] != NONE && in->job_gang[j] >= in->num_gangs) fail(ARMADA_E_INVALID, "job_gang out of range");
    if (in->job_node[j] == NONE && in->job_queue[j] == NONE) fail(ARMADA_E_INVALID, "queued job without queue");
  }
}

}  // namespace

struct ArmadaOracleNodeDb {
  NodeDb db;
  std::vector<JobCtx> jctx;
  explicit ArmadaOracleNodeDb(const ArmadaRoundInput* in) : db(in), jctx(in->num_jobs) {
    for (uint32_t j = 0; j < in->num_jobs; ++j) {
      jctx[j].job = j;
      jctx[j].row = in->class_static_row[in->job_class[j]];
    }
  }
};

#define ORACLE_GUARD(body)                 \
  try {                                    \
    body;                                  \
    return ARMADA_OK;                      \
  } catch (const Fail& f) {                \
    g_err = f.msg;                         \
    return f.code;                         \
  } catch (const std::exception& e) {      \
    g_err = e.what();                      \
    return ARMADA_E_INTERNAL;              \
  }

extern "C" {

const char* armada_oracle_last_error(void) { return g_err.c_str(); }

int32_t armada_oracle_round_schedule(const ArmadaRoundInput* in, ArmadaRoundOutput* out, ArmadaRoundStats* stats) {
  ORACLE_GUARD({
    validate(in);
    Round r(in);
    r.run(out);
    if (stats) *stats = r.stats;
  });
}

// QueueCandidateGangIteratorPQ.Less on two items given as {proposedQueueCost, currentQueueCost, queueBudget, itemSize,
// queue rank, priority-class priority} (test hook: TestQueueCandidateGangIteratorPQ_* vectors)
int32_t armada_oracle_pq_less(int32_t prioritise_larger, int32_t consider_priority, const double* a, const double* b) {
  Round::CandidateIt ci;
  ci.prioritise_larger = prioritise_larger != 0;
  ci.consider_priority = consider_priority != 0;
  Round::PQItem x{}, y{};
  x.proposed = a[0], x.current = a[1], x.budget = a[2], x.item_size = a[3], x.queue = (uint32_t)a[4], x.pc_priority = (int32_t)a[5];
  y.proposed = b[0], y.current = b[1], y.budget = b[2], y.item_size = b[3], y.queue = (uint32_t)b[4], y.pc_priority = (int32_t)b[5];
  return Round::pq_less(ci, x, y) ? 1 : 0;
}

double armada_oracle_drf_cost(uint32_t d, const int64_t* total, const double* multipliers, const int64_t* allocation) {
  return drf_unweighted((int)d, total, multipliers, allocation);
}

int32_t armada_oracle_nodedb_create(const ArmadaRoundInput* in, ArmadaOracleNodeDb** out) {
  ORACLE_GUARD({
    validate(in);
    auto* h = new ArmadaOracleNodeDb(in);
    try {
      h->db.bind_running_jobs();
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
  });
}

void armada_oracle_nodedb_destroy(ArmadaOracleNodeDb* db) { delete db; }

int32_t armada_oracle_nodedb_schedule_many(ArmadaOracleNodeDb* h, const uint32_t* jobs, uint32_t n, uint8_t* ok,
                                           uint32_t* node, int32_t* scheduled_at, int32_t* preempted_at,
                                           uint8_t* method) {
  ORACLE_GUARD({
    std::vector<JobCtx*> gang;
    for (uint32_t i = 0; i < n; ++i) {
      if (jobs[i] >= h->db.J) fail(ARMADA_E_INVALID, "job out of range");
      JobCtx& jc = h->jctx[jobs[i]];
      jc.has_pctx = false;
      jc.p_node = NONE;
      gang.push_back(&jc);
    }
    h->db.begin();
    bool all = h->db.schedule_many(gang);
    if (all) h->db.commit();
    else h->db.abort();
    if (ok) *ok = all ? 1 : 0;
    for (uint32_t i = 0; i < n; ++i) {
      JobCtx& jc = *gang[i];
      if (node) node[i] = jc.has_pctx ? jc.p_node : NONE;
      if (scheduled_at) scheduled_at[i] = jc.has_pctx ? jc.p_sched_at : ARMADA_NO_PRIORITY;
      if (preempted_at) preempted_at[i] = jc.has_pctx ? jc.p_preempted_at : ARMADA_NO_PRIORITY;
      if (method) method[i] = jc.has_pctx ? jc.p_method : (uint8_t)ARMADA_METHOD_NONE;
    }
  });
}

// The same inside a transaction that is always aborted: what SubmitChecker does per gang and
// executor (internal/scheduler/submitcheck.go:372-380) — the NodeDb is unchanged afterwards.
int32_t armada_oracle_nodedb_dry_run(ArmadaOracleNodeDb* h, const uint32_t* jobs, uint32_t n, uint8_t* ok, uint32_t* node) {
  ORACLE_GUARD({
    std::vector<JobCtx*> gang;
    for (uint32_t i = 0; i < n; ++i) {
      if (jobs[i] >= h->db.J) fail(ARMADA_E_INVALID, "job out of range");
      JobCtx& jc = h->jctx[jobs[i]];
      jc.has_pctx = false;
      jc.p_node = NONE;
      gang.push_back(&jc);
    }
    h->db.begin();
    bool all = h->db.schedule_many(gang);
    h->db.abort();
    if (ok) *ok = all ? 1 : 0;
    for (uint32_t i = 0; i < n; ++i)
      if (node) node[i] = (all && gang[i]->has_pctx) ? gang[i]->p_node : NONE;
  });
}

int32_t armada_oracle_nodedb_evict(ArmadaOracleNodeDb* h, uint32_t job) {
  ORACLE_GUARD({
    if (job >= h->db.J) fail(ARMADA_E_INVALID, "job out of range");
    h->db.evict(job);
  });
}

int32_t armada_oracle_nodedb_unbind(ArmadaOracleNodeDb* h, uint32_t job) {
  ORACLE_GUARD({
    if (job >= h->db.J) fail(ARMADA_E_INVALID, "job out of range");
    uint32_t n = h->db.bound_node[job];
    if (n != NONE) h->db.unbind(job, n);
    h->jctx[job].assigned_node = NONE;
    h->jctx[job].is_evicted = false;
    h->jctx[job].has_additional = false;
  });
}

int32_t armada_oracle_nodedb_add_evicted(ArmadaOracleNodeDb* h, uint32_t job, int32_t index) {
  ORACLE_GUARD({
    if (job >= h->db.J) fail(ARMADA_E_INVALID, "job out of range");
    uint32_t n = h->db.bound_node[job];
    if (n == NONE || !h->db.evicted_on_node[job]) fail(ARMADA_E_INVALID, "job is not evicted on a node");
    JobCtx& jc = h->jctx[job];
    jc.is_evicted = true;
    jc.assigned_node = n;
    jc.has_additional = true;
    if (index >= 0) {
      if (h->db.evicted_index_of_job[job] >= 0 || h->db.evicted_by_index.count(index))
        fail(ARMADA_E_INVALID, "duplicate evicted index");
      h->db.ev_insert(index, job);
      h->db.evicted_index_of_job[job] = index;
    }
  });
}

int32_t armada_oracle_nodedb_get_alloc(ArmadaOracleNodeDb* h, uint32_t node, int64_t* out) {
  ORACLE_GUARD({
    if (node >= h->db.N) fail(ARMADA_E_INVALID, "node out of range");
    for (int p = 0; p < h->db.PL; ++p)
      for (int d = 0; d < h->db.D; ++d) out[(size_t)p * h->db.D + d] = h->db.A(p, d, node);
  });
}

int32_t armada_oracle_nodedb_iterate(ArmadaOracleNodeDb* h, uint32_t row, int32_t priority,
                                     const int64_t* indexed_request, uint32_t* out_nodes, uint32_t cap,
                                     uint32_t* count) {
  ORACLE_GUARD({
    if (row >= h->db.in->num_static_rows) fail(ARMADA_E_INVALID, "row out of range");
    int p = h->db.level_of(priority);
    NodeDb::TypesIt m;
    h->db.types_it_init(m, row, p, indexed_request);
    uint32_t c = 0;
    for (uint32_t n = h->db.types_it_next(m); n != NONE; n = h->db.types_it_next(m)) {
      if (c < cap) out_nodes[c] = n;
      ++c;
    }
    *count = c;
  });
}

// NodeIndexKey / RoundedNodeIndexKeyFromResourceList + EncodeInt64/EncodeUint64, encoding.go:22-89
int32_t armada_oracle_node_index_key(uint32_t r, uint64_t node_type_id, const int64_t* quantities,
                                     const int64_t* resolution, uint64_t node_index, int rounded, uint8_t* out) {
  auto put = [&](uint64_t v) {
    for (int i = 7; i >= 0; --i) *out++ = (uint8_t)(v >> (8 * i));
  };
  put(node_type_id);
  for (uint32_t i = 0; i < r; ++i) {
    int64_t q = quantities[i];
    if (rounded) q = round_to_resolution(q, resolution[i]);
    put((uint64_t)q ^ 0x8000000000000000ull);
  }
  put(rounded ? node_index : 0);
  return ARMADA_OK;
}

uint32_t armada_oracle_abi_sizeof(uint32_t which) {
  switch (which) {
    case 0: return (uint32_t)sizeof(ArmadaRoundInput);
    case 1: return (uint32_t)sizeof(ArmadaRoundOutput);
    case 2: return (uint32_t)sizeof(ArmadaRoundStats);
    default: return 0;
  }
}

}  // extern "C"
