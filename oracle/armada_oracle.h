/*
 * armada_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A single-threaded C++17 restatement of the reference's scheduling round
 * (armadaproject/armada internal/scheduler/{nodedb,scheduling,internaltypes,jobdb}),
 * consuming the same ArmadaRoundInput / ArmadaRoundOutput structs as the product
 * library (include/armada_b200.h).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this library.
 *
 * Parity status: PINNED against the reference's own known-answer tests transcribed in
 * tests/test_golden_*.py (nodedb_test.go, nodeiteration_test.go, gang_scheduler_test.go,
 * queue_scheduler_test.go, preempting_queue_scheduler_test.go, context/scheduling_test.go,
 * fairness_test.go).  The Go toolchain is absent in the build container, so the
 * reference itself could not be executed here.
 */
#ifndef ARMADA_ORACLE_H_
#define ARMADA_ORACLE_H_

#include "armada_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Whole round: PreemptingQueueScheduler.Schedule (preempting_queue_scheduler.go:84-285). */
int32_t armada_oracle_round_schedule(const ArmadaRoundInput* in, ArmadaRoundOutput* out,
                                     ArmadaRoundStats* stats);
const char* armada_oracle_last_error(void);

/* DominantResourceFairness.UnweightedCostFromAllocation (fairness/fairness.go:103-105). */
/* QueueCandidateGangIteratorPQ.Less (queue_scheduler.go:628-674) on {proposed, current, budget, itemSize, queue rank, pc priority} */
int32_t armada_oracle_pq_less(int32_t prioritise_larger, int32_t consider_priority, const double* a, const double* b);
double armada_oracle_drf_cost(uint32_t d, const int64_t* total, const double* multipliers,
                              const int64_t* allocation);

/* ---- NodeDb-level surface for the nodedb_test.go / nodeiteration_test.go vectors ---- */
typedef struct ArmadaOracleNodeDb ArmadaOracleNodeDb;
/* NewNodeDb + CreateAndInsertWithJobDbJobsWithTxn for every node (running jobs in `in`
 * are bound at their scheduled-at priority, nodedb.go:43-60). */
int32_t armada_oracle_nodedb_create(const ArmadaRoundInput* in, ArmadaOracleNodeDb** out);
void armada_oracle_nodedb_destroy(ArmadaOracleNodeDb* db);
/* ScheduleManyWithTxn over `jobs` inside one txn, committed iff all fit
 * (nodedb.go:386-418 + gang_scheduler.go:226-237).  Per-job results are written even on
 * failure (node = ARMADA_NONE for members that did not fit / were not attempted). */
int32_t armada_oracle_nodedb_schedule_many(ArmadaOracleNodeDb* db, const uint32_t* jobs, uint32_t n,
                                           uint8_t* ok, uint32_t* node, int32_t* scheduled_at,
                                           int32_t* preempted_at, uint8_t* method);
/* EvictJobsFromNode / UnbindJobFromNode + Upsert (nodedb.go:960-1003,1039-1098). */
int32_t armada_oracle_nodedb_evict(ArmadaOracleNodeDb* db, uint32_t job);
int32_t armada_oracle_nodedb_unbind(ArmadaOracleNodeDb* db, uint32_t job);
/* ScheduleManyWithTxn inside a transaction that is always aborted (SubmitChecker,
 * submitcheck.go:372-380): the NodeDb is unchanged afterwards. */
int32_t armada_oracle_nodedb_dry_run(ArmadaOracleNodeDb* db, const uint32_t* jobs, uint32_t n, uint8_t* ok, uint32_t* node);
/* Mark an evicted job as pinned to its node for re-scheduling (jctx.SetAssignedNode,
 * eviction.go:245-252) and register it for fair preemption
 * (AddEvictedJobSchedulingContextWithTxn, nodedb.go:1193-1203). */
int32_t armada_oracle_nodedb_add_evicted(ArmadaOracleNodeDb* db, uint32_t job, int32_t index);
/* AllocatableByPriority of one node, out[PL][D]. */
int32_t armada_oracle_nodedb_get_alloc(ArmadaOracleNodeDb* db, uint32_t node, int64_t* out);
/* NodeTypesIterator visiting order (nodeiteration.go:74-208): node types = those with
 * type_match[row] set; request = indexed-resource requests [R]; returns count. */
int32_t armada_oracle_nodedb_iterate(ArmadaOracleNodeDb* db, uint32_t row, int32_t priority,
                                     const int64_t* indexed_request, uint32_t* out_nodes,
                                     uint32_t cap, uint32_t* count);
/* nodeDbKey / RoundedNodeIndexKeyFromResourceList bytes (encoding.go:37-89): writes
 * 8*(R+2) bytes. */
int32_t armada_oracle_node_index_key(uint32_t r, uint64_t node_type_id, const int64_t* quantities,
                                     const int64_t* resolution, uint64_t node_index, int rounded,
                                     uint8_t* out);

/* sizeof() of ArmadaRoundInput / Output / Stats as compiled into the oracle (0, 1, 2). */
uint32_t armada_oracle_abi_sizeof(uint32_t which);

#ifdef __cplusplus
}
#endif
#endif
