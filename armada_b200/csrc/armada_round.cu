// armada_round.cu — hand-written CUDA (sm_100a) for Armada's scheduling round.
//
// Reference path being replaced (relative to internal/scheduler/ of armadaproject/armada):
//   scheduling/preempting_queue_scheduler.go:84-285   PreemptingQueueScheduler.Schedule
//   scheduling/queue_scheduler.go:91-272,315-383,495-674  QueueScheduler loop, gang iterator, cost PQ
//   scheduling/gang_scheduler.go:100-254              GangScheduler
//   nodedb/nodedb.go:386-512,605-1098                 ScheduleMany / SelectNode cascade / bind / evict / unbind
//   nodedb/nodeiteration.go:74-382 + encoding.go      best-fit ordered walk
//   scheduling/eviction.go:132-273                    evictors
//   scheduling/context/{scheduling,queue}.go          per-round accounting, fair shares
//   scheduling/fairness/fairness.go:99-105            DRF cost
//
// B200 mapping:
//   * data-parallel phases (bind running jobs, both evictors, gang completion, accounting,
//     result algebra) are grid-wide kernels over the job SoA with int64 atomics on the node SoA;
//   * evicted jobs are ranked per queue with an LSD radix sort on (queue, order-rank) keys;
//   * the inherently sequential QueueScheduler loop runs in ONE persistent CTA: warp 0 is the
//     control warp (DRF arg-min over queues by warp reduction, gang iterator, constraints,
//     bind), warps 0..31 jointly maintain per-(job shape, priority level) best-fit tournament
//     trees in shared memory whose leaves are coalesced 128-node tile scans of the HBM/L2
//     resident node SoA with redux.sync arg-min on a packed 64-bit (resources…, node-id) key.
//     A probe is a root read; a bind re-scans one tile per affected tree.
//   Tensor cores are not used: this is integer compare/select work.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "armada_dev.h"

#define NONE ARMADA_NONE
#define NOPRIO ARMADA_NO_PRIORITY
#define KEY_INF 0xFFFFFFFFFFFFFFFFull
#define FULL 0xFFFFFFFFu

namespace {

thread_local std::string g_last_error;

// =====================================================================================
// device helpers
// =====================================================================================
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k) {
  unsigned hi = (unsigned)(k >> 32);
  unsigned mhi = __reduce_min_sync(FULL, hi);
  unsigned lo = (hi == mhi) ? (unsigned)k : 0xFFFFFFFFu;
  unsigned mlo = __reduce_min_sync(FULL, lo);
  return ((unsigned long long)mhi << 32) | mlo;
}

__device__ __forceinline__ void atomic_add_i64(int64_t* p, int64_t v) {
  atomicAdd((unsigned long long*)p, (unsigned long long)v);
}

// priorityCutoffFor (nodedb.go:1031-1036): preemptible ⇒ scheduled priority, else every level.
__device__ __forceinline__ int32_t cutoff_for(const DevCfg& c, uint32_t pc, int32_t scheduled_priority) {
  return c.pcs[pc].preemptible ? scheduled_priority : INT32_MAX;
}

// =====================================================================================
// grid-wide kernels (data-parallel phases)
// =====================================================================================

// newAllocatableByPriorityAndResourceType (nodedb.go:1185-1191) + per-job state reset.
__global__ void k_reset(DevCfg c, DevPtrs P) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t DN = (size_t)c.D * c.N;
  for (size_t k = i; k < (size_t)c.PL * DN; k += stride) P.alloc[k] = P.node_allocatable[k % DN];
  for (size_t j = i; j < c.J; j += stride) {
    P.bound_node[j] = NONE;
    P.evicted_on_node[j] = 0;
    P.sched_prio[j] = NOPRIO;
    P.q_successful[j] = P.q_unsuccessful[j] = P.q_rescheduled[j] = P.q_evicted[j] = P.q_reason[j] = 0;
    P.jc_evicted[j] = 0;
    uint32_t g = P.job_gang[j];
    P.jc_card[j] = g == NONE ? 1u : P.gang_card[g];
    P.assigned_node[j] = NONE;
    P.in_preempted[j] = P.in_scheduled[j] = P.in_sae[j] = P.ev_mark[j] = 0;
    P.has_pctx[j] = 0;
    P.res_node[j] = NONE;
    P.res_sched_at[j] = NOPRIO;
    P.res_preempted_at[j] = NOPRIO;
    P.res_method[j] = 0;
    P.ev_index_of_job[j] = -1;
    P.ev_alive[j] = 0;
  }
  for (size_t g = i; g < c.G; g += stride) {
    P.gang_fill[g] = 0;
    P.gang_ev[g] = 0;
  }
  for (size_t n = i; n < c.N; n += stride) P.fp_epoch[n] = 0;
  for (size_t k = i; k < (size_t)c.C; k += stride) P.unfeasible[k] = 0;
  if (i < 16) P.counters[i] = 0;
  if (i < 8) P.stats[i] = 0;
}

// CreateAndInsertWithJobDbJobsWithTxn (nodedb.go:43-60): bind every running job at its
// scheduled-at priority.  Integer atomics commute, so the result is order independent.
__global__ void k_bind_running(DevCfg c, DevPtrs P) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= c.J) return;
  uint32_t n = P.job_node0[j];
  if (n == NONE) return;
  uint32_t cls = P.job_class[j];
  uint32_t pc = P.class_pc[cls];
  int32_t pr = P.job_sap0[j];
  if (pr == NOPRIO) pr = c.pcs[pc].priority;
  int32_t cut = cutoff_for(c, pc, pr);
  for (int p = 0; p < c.PL; ++p) {
    if (c.priorities[p] > cut) continue;
    for (int d = 0; d < c.D; ++d) {
      int64_t r = P.class_req_node[(size_t)cls * c.D + d];
      if (r) atomic_add_i64(&P.alloc[((size_t)p * c.D + d) * c.N + n], -r);
    }
  }
  P.bound_node[j] = n;
  P.sched_prio[j] = pr;
}

// DominantResourceFairness.UnweightedCostFromAllocation (fairness/fairness.go:103-105).
__device__ double drf_cost_seq(const DevCfg& c, const int64_t* a) {
  double result = -INFINITY;
  for (int d = 0; d < c.D; ++d) {
    double frac = 0.0;
    if (c.total_resources[d] != 0) frac = __ddiv_rn((double)a[d], (double)c.total_resources[d]);
    double v = __dmul_rn(frac, c.drf_mult[d]);
    if (v > result) result = v;
  }
  if (result != result) return result;
  return result > 0.0 ? result : 0.0;
}

// AddQueueSchedulingContext + UpdateFairShares (context/scheduling.go:104-156,252-332).  Q is
// small; one thread restates the iterative water-filling exactly (same op order ⇒ same doubles).
__global__ void k_queue_init(DevCfg c, DevPtrs P) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int Q = c.Q, D = c.D;
  double weight_sum = 0.0;
  for (int q = 0; q < Q; ++q) {
    for (int d = 0; d < D; ++d) {
      int64_t s = 0;
      for (int pc = 0; pc < c.PC; ++pc) {
        int64_t v = P.queue_alloc_pc0[((size_t)q * c.PC + pc) * D + d];
        P.q_alloc_pc[((size_t)q * c.PC + pc) * D + d] = v;
        s += v;
      }
      P.q_alloc[(size_t)q * D + d] = s;
    }
    P.q_tokens[q] = P.queue_tokens0[q];
    weight_sum += P.queue_weight[q];
  }
  for (int d = 0; d < D; ++d) P.s_scheduled[d] = P.s_evicted[d] = 0;
  for (int k = 0; k < 8; ++k) P.s_counts[k] = 0;
  P.s_counts[4] = __double_as_longlong(c.global_tokens);
  bool total_all_zero = true;
  for (int d = 0; d < D; ++d)
    if (c.total_resources[d] != 0) total_all_zero = false;
  // scratch in q_fair: [q][0]=fair, [1]=capped, [2]=uncapped ; cds/spare/achieved in registers per pass
  // Q <= ARMADA_DEV_MAX_QUEUES, keep small local arrays
  double cds[ARMADA_DEV_MAX_QUEUES], spare[ARMADA_DEV_MAX_QUEUES];
  bool achieved[ARMADA_DEV_MAX_QUEUES];
  for (int q = 0; q < Q; ++q) {
    double v = 1.0;
    if (!total_all_zero) v = drf_cost_seq(c, P.queue_cdemand + (size_t)q * D);
    cds[q] = v;
    spare[q] = 0.0;
    achieved[q] = false;
    P.q_fair[q * 3 + 0] = __ddiv_rn(P.queue_weight[q], weight_sum);
    P.q_fair[q * 3 + 1] = 0.0;
    P.q_fair[q * 3 + 2] = 0.0;
  }
  double unallocated = 1.0;
  for (int it = 0; it < 10 && unallocated > 0.01; ++it) {
    double total_weight = 0.0;
    for (int q = 0; q < Q; ++q)
      if (!achieved[q]) total_weight += P.queue_weight[q];
    for (int q = 0; q < Q; ++q) {
      double tw = total_weight;
      if (achieved[q]) tw += P.queue_weight[q];
      double share = __dmul_rn(__ddiv_rn(P.queue_weight[q], tw), __dsub_rn(unallocated, spare[q]));
      P.q_fair[q * 3 + 2] = __dadd_rn(P.q_fair[q * 3 + 2], share);
    }
    if (total_weight <= 0.0) break;
    for (int q = 0; q < Q; ++q)
      if (!achieved[q])
        P.q_fair[q * 3 + 1] = __dadd_rn(P.q_fair[q * 3 + 1], __dmul_rn(__ddiv_rn(P.queue_weight[q], total_weight), unallocated));
    unallocated = 0.0;
    for (int q = 0; q < Q; ++q) {
      double sp = __dsub_rn(P.q_fair[q * 3 + 1], cds[q]);
      if (sp > 0) {
        P.q_fair[q * 3 + 1] = cds[q];
        achieved[q] = true;
        spare[q] = sp;
        unallocated = __dadd_rn(unallocated, sp);
      } else {
        spare[q] = 0;
      }
    }
  }
}

// evictJobFromNodeInPlace (nodedb.go:974-1003) applied with atomics: markAllocatable at the
// job's cutoff, markAllocated at EvictedPriority ⇒ levels 1.. (p>=0, p<=cutoff) += req.
__device__ void evict_alloc(const DevCfg& c, const DevPtrs& P, uint32_t j, uint32_t n) {
  uint32_t cls = P.job_class[j];
  int32_t cut = cutoff_for(c, P.class_pc[cls], P.sched_prio[j]);
  for (int p = 1; p < c.PL; ++p) {  // level 0 (EvictedPriority) nets to zero
    if (c.priorities[p] > cut) continue;
    for (int d = 0; d < c.D; ++d) {
      int64_t r = P.class_req_node[(size_t)cls * c.D + d];
      if (r) atomic_add_i64(&P.alloc[((size_t)p * c.D + d) * c.N + n], r);
    }
  }
}

// NodeEvictor + balancing job filter (eviction.go:85-103, preempting_queue_scheduler.go:99-134).
// The filter only depends on per-queue constants (queue allocation is read BEFORE any eviction is
// applied to the scheduling context), so it is evaluated once per queue into evict_queue[].
__global__ void k_evict_queue_flags(DevCfg c, DevPtrs P, uint8_t* evict_queue) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= c.Q) return;
  double actual = drf_cost_seq(c, P.q_alloc + (size_t)q * c.D);
  double dc = P.q_fair[q * 3 + 1], fs = P.q_fair[q * 3 + 0];
  double fair = (dc != dc || fs != fs) ? NAN : fmax(dc, fs);  // math.Max
  if (c.protect_uncapped) fair = P.q_fair[q * 3 + 2];
  double fraction = __ddiv_rn(actual, fair);
  evict_queue[q] = !(fraction <= c.protected_fraction);
}

__global__ void k_evict_balance(DevCfg c, DevPtrs P, const uint8_t* evict_queue) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= c.J) return;
  uint32_t n = P.bound_node[j];
  if (n == NONE || P.evicted_on_node[j]) return;
  uint32_t q = P.job_queue[j];
  if (q == NONE) return;                                        // invalid_queue
  if (!c.pcs[P.class_pc[P.job_class[j]]].preemptible) return;   // job_not_preemptible
  if (!evict_queue[q]) return;                                  // below_protected_fair_share
  P.ev_mark[j] = 1;
  P.evicted_on_node[j] = 1;
  evict_alloc(c, P, j, n);
  uint32_t g = P.job_gang[j];
  if (g != NONE) P.gang_ev[g] = 1;
}

// NewOversubscribedEvictor (eviction.go:132-184): a node is processed iff some p >= 0 has a
// negative allocatable; a job is evicted iff preemptible and its scheduled-at level is one.
// oversub_mask[n] is computed first (k_oversub_mask) because evicting changes alloc.
__global__ void k_oversub_mask(DevCfg c, DevPtrs P, uint32_t* mask) {
  size_t n = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (n >= c.N) return;
  uint32_t m = 0;
  for (int p = 1; p < c.PL; ++p)
    for (int d = 0; d < c.D; ++d)
      if (P.alloc[((size_t)p * c.D + d) * c.N + n] < 0) m |= 1u << p;
  mask[n] = m;
}

__global__ void k_evict_oversub(DevCfg c, DevPtrs P, const uint32_t* mask) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= c.J) return;
  uint32_t n = P.bound_node[j];
  if (n == NONE || P.evicted_on_node[j]) return;
  if (P.job_queue[j] == NONE) return;
  if (!c.pcs[P.class_pc[P.job_class[j]]].preemptible) return;
  int32_t pr = P.sched_prio[j];
  if (pr == NOPRIO) return;
  uint32_t m = mask[n];
  if (!m) return;
  int lvl = -1;
  for (int p = 0; p < c.PL; ++p)
    if (c.priorities[p] == pr) lvl = p;
  if (lvl < 0 || !((m >> lvl) & 1u)) return;
  P.ev_mark[j] = 1;
  P.evicted_on_node[j] = 1;
  evict_alloc(c, P, j, n);
  uint32_t g = P.job_gang[j];
  if (g != NONE) P.gang_ev[g] = 1;
}

// evictGangs (preempting_queue_scheduler.go:353-420): every bound, not yet evicted member of a
// gang with an evicted member is evicted too (its own node is in gangNodeIds by construction).
__global__ void k_evict_gangs(DevCfg c, DevPtrs P) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= c.J) return;
  uint32_t g = P.job_gang[j];
  if (g == NONE || !P.gang_ev[g]) return;
  uint32_t n = P.bound_node[j];
  if (n == NONE || P.evicted_on_node[j]) return;
  if (P.sched_prio[j] == NOPRIO) return;
  P.ev_mark[j] = 1;
  P.evicted_on_node[j] = 1;
  evict_alloc(c, P, j, n);
}

// sctx.EvictJob for every evicted job (context/scheduling.go:470-491, queue.go:296-331),
// setEvictedGangCardinality (:458-479), result-set algebra of :141-194, and emission of the
// (queue, order-rank) sort key for InMemoryJobRepository.sortQueue (jobiteration.go:72-95).
__global__ void k_evict_account(DevCfg c, DevPtrs P, int pass) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= c.J) return;
  if (!P.ev_mark[j]) return;
  P.ev_mark[j] = 0;
  uint32_t q = P.job_queue[j];
  uint32_t cls = P.job_class[j];
  uint32_t pc = P.class_pc[cls];
  bool sched = P.q_successful[j], resched = P.q_rescheduled[j];
  if (sched || resched) {
    P.q_successful[j] = 0;
    P.q_rescheduled[j] = 0;
  } else {
    P.q_evicted[j] = 1;
  }
  for (int d = 0; d < c.D; ++d) {
    int64_t r = P.class_req_raw[(size_t)cls * c.D + d];
    if (!r) continue;
    atomic_add_i64(&P.q_alloc[(size_t)q * c.D + d], -r);
    atomic_add_i64(&P.q_alloc_pc[((size_t)q * c.PC + pc) * c.D + d], -r);
    if (sched) atomic_add_i64(&P.s_scheduled[d], -r);
    else atomic_add_i64(&P.s_evicted[d], r);
  }
  if (sched) atomic_add_i64(&P.s_counts[0], -1);
  else atomic_add_i64(&P.s_counts[2], 1);
  // jctx for re-scheduling (eviction.go:241-256)
  P.jc_evicted[j] = 1;
  P.assigned_node[j] = P.bound_node[j];
  uint32_t g = P.job_gang[j];
  P.jc_card[j] = g == NONE ? 1u : P.gang_count[g];
  P.has_pctx[j] = 0;
  if (pass == 1) {
    P.in_preempted[j] = 1;
  } else {
    if (P.in_scheduled[j]) {
      P.in_scheduled[j] = 0;
      P.in_sae[j] = 1;
    } else {
      P.in_preempted[j] = 1;
    }
  }
  uint32_t slot = atomicAdd(&P.counters[0], 1u);
  P.sort_keys[slot] = ((unsigned long long)q << 32) | P.job_rank[j];
  P.sort_vals[slot] = (uint32_t)j;
}

__global__ void k_clear_gang_ev(DevCfg c, DevPtrs P) {
  size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (g < c.G) P.gang_ev[g] = 0;
}

// ---- LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass --------------------------
// Ranks eviction candidates per queue: key = queue << 32 | SchedulingOrderCompare rank.
#define RS_THREADS 256
#define RS_ITEMS 16
__global__ void k_rs_hist(const unsigned long long* keys, uint32_t n, int shift, uint32_t* hist /*[256][blocks]*/) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * RS_THREADS * RS_ITEMS;
  for (int i = 0; i < RS_ITEMS; ++i) {
    size_t k = base + (size_t)i * RS_THREADS + threadIdx.x;
    if (k < n) atomicAdd(&h[(keys[k] >> shift) & 255], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}
// exclusive scan over the digit-major histogram (single block; length = 256 * blocks)
__global__ void k_rs_scan(uint32_t* hist, uint32_t len) {
  __shared__ uint32_t carry;
  __shared__ uint32_t tmp[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < len; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < len ? hist[i] : 0;
    tmp[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      uint32_t t = threadIdx.x >= (unsigned)off ? tmp[threadIdx.x - off] : 0;
      __syncthreads();
      tmp[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t incl = tmp[threadIdx.x];
    if (i < len) hist[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
}
__global__ void k_rs_scatter(const unsigned long long* keys, const uint32_t* vals, unsigned long long* keys_out,
                             uint32_t* vals_out, uint32_t n, int shift, const uint32_t* hist) {
  // stable within a block: one thread per digit value walks the block's items in order
  __shared__ unsigned long long sk[RS_THREADS * RS_ITEMS];
  __shared__ uint32_t sv[RS_THREADS * RS_ITEMS];
  size_t base = (size_t)blockIdx.x * RS_THREADS * RS_ITEMS;
  uint32_t cnt = 0;
  for (int i = 0; i < RS_ITEMS; ++i) {
    size_t k = base + (size_t)i * RS_THREADS + threadIdx.x;
    uint32_t li = i * RS_THREADS + threadIdx.x;
    if (k < n) {
      sk[li] = keys[k];
      sv[li] = vals[k];
    }
  }
  __syncthreads();
  uint32_t m = n - base < (size_t)RS_THREADS * RS_ITEMS ? (uint32_t)(n - base) : RS_THREADS * RS_ITEMS;
  uint32_t off = hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
  for (uint32_t li = 0; li < m; ++li) {
    if (((sk[li] >> shift) & 255) == threadIdx.x) {
      keys_out[off + cnt] = sk[li];
      vals_out[off + cnt] = sv[li];
      ++cnt;
    }
  }
}

// evq_start[q] = first position in the sorted evicted list whose queue is >= q.
__global__ void k_evq_bounds(DevCfg c, DevPtrs P, const unsigned long long* keys, uint32_t n) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > c.Q) return;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if ((uint32_t)(keys[mid] >> 32) < (uint32_t)q) lo = mid + 1;
    else hi = mid;
  }
  P.evq_start[q] = lo;
}
__global__ void k_copy_u32(uint32_t* dst, const uint32_t* src, uint32_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// =====================================================================================
// persistent schedule-pass kernel (one CTA)
// =====================================================================================
enum { CMD_EXIT = 0, CMD_REFRESH = 1, CMD_BUILD = 2, CMD_SCAN = 3 };

struct QItem {  // QueueCandidateGangIteratorItem (queue_scheduler.go:601-626)
  double proposed, current, budget, size;
  int32_t pcprio;
  int32_t q;  // -1 = absent
};

struct PassArgs {
  int with_queued;        // queued-jobs iterator present (pass 1)
  int skip_key_check;     // skipUnsuccessfulSchedulingKeyCheck
  int consider_priority;  // considerPriorityClassPriority
  int assign_indices;     // run addEvictedJobsToNodeDb first
  int pass;               // 1 or 2
  uint32_t num_evicted;   // evicted jobs in evq_jobs
};

struct SmemHdr {
  int cmd;
  int cmd_slot;
  int cmd_level;
  uint32_t cmd_node;
  uint32_t cmd_levelmask;
  uint32_t cmd_row;
  int64_t cmd_req[ARMADA_MAX_RESOURCES];
  unsigned long long cmd_result;
  unsigned long long partial[32];
  // slots
  int num_active_slots;
  int slot_level[ARMADA_DEV_MAX_SLOTS];
  uint32_t slot_row[ARMADA_DEV_MAX_SLOTS];
  int64_t slot_req[ARMADA_DEV_MAX_SLOTS][ARMADA_MAX_RESOURCES];
  // queue state
  int64_t q_alloc[ARMADA_DEV_MAX_QUEUES][ARMADA_MAX_RESOURCES];
  int64_t ng_total[ARMADA_DEV_MAX_QUEUES][ARMADA_MAX_RESOURCES];  // next gang TotalResourceRequests
  double q_proposed[ARMADA_DEV_MAX_QUEUES], q_current[ARMADA_DEV_MAX_QUEUES], q_size[ARMADA_DEV_MAX_QUEUES],
      q_budget[ARMADA_DEV_MAX_QUEUES], q_tokens[ARMADA_DEV_MAX_QUEUES];
  int32_t q_pcprio[ARMADA_DEV_MAX_QUEUES];
  uint32_t ng_first[ARMADA_DEV_MAX_QUEUES];   // single job id, or gang id when ng_is_gang
  uint32_t ng_count[ARMADA_DEV_MAX_QUEUES];
  uint32_t ev_cur[ARMADA_DEV_MAX_QUEUES], ev_end[ARMADA_DEV_MAX_QUEUES];
  uint32_t qd_cur[ARMADA_DEV_MAX_QUEUES], qd_end[ARMADA_DEV_MAX_QUEUES];
  uint32_t jobs_seen[ARMADA_DEV_MAX_QUEUES];
  uint8_t q_in_pq[ARMADA_DEV_MAX_QUEUES];       // item currently in the heap
  uint8_t ng_is_gang[ARMADA_DEV_MAX_QUEUES];
  uint8_t ng_all_evicted[ARMADA_DEV_MAX_QUEUES];
  uint8_t it_only_evicted[ARMADA_DEV_MAX_QUEUES];   // QueuedGangIterator.onlyYieldEvicted
  uint8_t it_has_next[ARMADA_DEV_MAX_QUEUES];
  uint8_t q_only_evicted_for_queue[ARMADA_DEV_MAX_QUEUES];
  // sctx
  int64_t s_scheduled[ARMADA_MAX_RESOURCES], s_evicted[ARMADA_MAX_RESOURCES];
  long long num_sched_jobs, num_sched_gangs, num_evicted_jobs;
  double global_tokens;
  uint32_t num_unfeasible;
  int only_yield_evicted;  // CostBasedCandidateGangIterator.onlyYieldEvicted
  uint32_t undo_n;
  uint32_t fp_epoch;
  unsigned long long st_iters, st_probes, st_placements, st_fair, st_rescans;
  int error;
};

struct Trees {  // views into dynamic shared memory
  unsigned long long* root;  // [slots]
  unsigned long long* l2;    // [slots][num_groups]
  unsigned long long* leaf;  // [slots][num_tiles]
};

__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 1, 1024;" ::: "memory"); }

// Packed best-fit key of node n at level p for a request (encoding.go:37-58 bit-packed): feasible
// nodes only (static ok ∧ alloc[p][d] >= req[d] ∀d, nodematching.go:147-197) else KEY_INF.
__device__ __forceinline__ unsigned long long node_key(const DevCfg& c, const int64_t* alloc_p, uint32_t n,
                                                       const int64_t* req, bool static_ok) {
  bool ok = static_ok;
  unsigned long long key = n;
#pragma unroll
  for (int d = 0; d < ARMADA_MAX_RESOURCES; ++d) {
    if (d < c.D) {
      int64_t a = alloc_p[(size_t)d * c.N + n];
      ok = ok && a >= req[d];
    }
  }
  if (!ok) return KEY_INF;
  for (int i = 0; i < c.R; ++i) {
    int64_t a = alloc_p[(size_t)c.indexed_resource[i] * c.N + n];
    key |= (unsigned long long)a << c.key_shift[i];
  }
  return key;
}

// StaticJobRequirementsMet (nodematching.go:161-190): bitmap row × node static class, and
// total resources >= request on all D.
__device__ __forceinline__ bool static_ok_inline(const DevCfg& c, const DevPtrs& P, uint32_t row, const int64_t* req, uint32_t n) {
  uint32_t sc = P.node_sclass[n];
  bool ok = (P.static_match[(size_t)row * c.sw + (sc >> 5)] >> (sc & 31)) & 1u;
  for (int d = 0; d < c.D; ++d) ok = ok && req[d] <= P.node_total[(size_t)d * c.N + n];
  return ok;
}

// Warp-collective: min key over the nodes of one leaf tile.
__device__ unsigned long long tile_min(const DevCfg& c, const DevPtrs& P, int level, const int64_t* req, const uint32_t* static_words,
                                       uint32_t row, uint32_t tile) {
  const int64_t* alloc_p = P.alloc + (size_t)level * c.D * c.N;
  uint32_t base = tile << c.tile_shift;
  int per_lane = 1 << (c.tile_shift - 5);
  unsigned long long best = KEY_INF;
  for (int k = 0; k < per_lane; ++k) {
    uint32_t n = base + (k << 5) + lane_id();
    if (n < c.N) {
      bool sok;
      if (static_words) sok = (static_words[(base >> 5) + k] >> lane_id()) & 1u;
      else sok = static_ok_inline(c, P, row, req, n);
      unsigned long long key = node_key(c, alloc_p, n, req, sok);
      best = key < best ? key : best;
    }
  }
  return warp_min_u64(best);
}

// Rebuild the root path of `tile` for slot s (warp-collective).
__device__ void refresh_path(const DevCfg& c, const DevPtrs& P, SmemHdr* H, Trees T, int s, uint32_t tile) {
  const uint32_t* sw = P.slot_static + (size_t)s * ((c.N + 31) / 32);
  unsigned long long m = tile_min(c, P, H->slot_level[s], H->slot_req[s], sw, 0, tile);
  if (lane_id() == 0) T.leaf[(size_t)s * c.num_tiles + tile] = m;
  __syncwarp();
  uint32_t g = tile >> 5;
  uint32_t t = (g << 5) + lane_id();
  unsigned long long v = t < (uint32_t)c.num_tiles ? T.leaf[(size_t)s * c.num_tiles + t] : KEY_INF;
  v = warp_min_u64(v);
  if (lane_id() == 0) T.l2[(size_t)s * c.num_groups + g] = v;
  __syncwarp();
  unsigned long long r = KEY_INF;
  for (int gi = lane_id(); gi < c.num_groups; gi += 32) {
    unsigned long long x = T.l2[(size_t)s * c.num_groups + gi];
    r = x < r ? x : r;
  }
  r = warp_min_u64(r);
  if (lane_id() == 0) T.root[s] = r;
  __syncwarp();
}

__device__ void do_work(const DevCfg& c, const DevPtrs& P, SmemHdr* H, Trees T) {
  int cmd = H->cmd;
  int w = warp_id();
  if (cmd == CMD_REFRESH) {
    uint32_t tile = H->cmd_node >> c.tile_shift;
    for (int s = w; s < H->num_active_slots; s += 32) {
      if (!((H->cmd_levelmask >> H->slot_level[s]) & 1u)) continue;
      refresh_path(c, P, H, T, s, tile);
      if (lane_id() == 0) atomicAdd(&H->st_rescans, 1ull);
    }
  } else if (cmd == CMD_BUILD) {
    int s = H->cmd_slot;
    uint32_t* sw = P.slot_static + (size_t)s * ((c.N + 31) / 32);
    // static-ok bitmap
    for (uint32_t word = w; word < (c.N + 31) / 32; word += 32) {
      uint32_t n = (word << 5) + lane_id();
      bool ok = n < c.N && static_ok_inline(c, P, H->slot_row[s], H->slot_req[s], n);
      uint32_t b = __ballot_sync(FULL, ok);
      if (lane_id() == 0) sw[word] = b;
    }
    __threadfence_block();
    bar_all();
    for (uint32_t tile = w; tile < (uint32_t)c.num_tiles; tile += 32) {
      unsigned long long m = tile_min(c, P, H->slot_level[s], H->slot_req[s], sw, 0, tile);
      if (lane_id() == 0) T.leaf[(size_t)s * c.num_tiles + tile] = m;
    }
    bar_all();
    for (uint32_t g = w; g < (uint32_t)c.num_groups; g += 32) {
      uint32_t t = (g << 5) + lane_id();
      unsigned long long v = t < (uint32_t)c.num_tiles ? T.leaf[(size_t)s * c.num_tiles + t] : KEY_INF;
      v = warp_min_u64(v);
      if (lane_id() == 0) T.l2[(size_t)s * c.num_groups + g] = v;
    }
    bar_all();
    if (w == 0) {
      unsigned long long r = KEY_INF;
      for (int gi = lane_id(); gi < c.num_groups; gi += 32) {
        unsigned long long x = T.l2[(size_t)s * c.num_groups + gi];
        r = x < r ? x : r;
      }
      r = warp_min_u64(r);
      if (lane_id() == 0) T.root[s] = r;
    }
  } else if (cmd == CMD_SCAN) {
    unsigned long long best = KEY_INF;
    for (uint32_t tile = w; tile < (uint32_t)c.num_tiles; tile += 32) {
      unsigned long long m = tile_min(c, P, H->cmd_level, H->cmd_req, nullptr, H->cmd_row, tile);
      best = m < best ? m : best;
    }
    if (lane_id() == 0) H->partial[w] = best;
    bar_all();
    if (w == 0) {
      unsigned long long v = H->partial[lane_id()];
      v = warp_min_u64(v);
      if (lane_id() == 0) H->cmd_result = v;
    }
  }
}

// ---- control-warp helpers (all 32 lanes execute these uniformly) -----------------------------
struct Ctl {
  const DevCfg& c;
  const DevPtrs& P;
  SmemHdr* H;
  Trees T;
  const PassArgs& a;

  __device__ void issue(int cmd) {
    __syncwarp();
    if (lane_id() == 0) H->cmd = cmd;
    __syncwarp();
    bar_all();
    do_work(c, P, H, T);
    bar_all();
  }
  __device__ void refresh(uint32_t n, uint32_t levelmask) {
    if (H->num_active_slots == 0) return;
    if (lane_id() == 0) {
      H->cmd_node = n;
      H->cmd_levelmask = levelmask;
    }
    __threadfence_block();
    issue(CMD_REFRESH);
  }
  __device__ int level_of(int32_t priority) const {
    for (int p = 0; p < c.PL; ++p)
      if (c.priorities[p] == priority) return p;
    return -1;
  }
  __device__ uint32_t levelmask_upto(int32_t cutoff) const {  // levels with priority <= cutoff
    uint32_t m = 0;
    for (int p = 0; p < c.PL; ++p)
      if (c.priorities[p] <= cutoff) m |= 1u << p;
    return m;
  }

  // DRF cost, lane d computes resource d (fairness.go:99-105); exact max ⇒ order independent.
  __device__ double drf_cost(const int64_t* a /*smem or global, [D]*/) const {
    double v = -INFINITY;
    int d = lane_id();
    if (d < c.D) {
      double frac = 0.0;
      if (c.total_resources[d] != 0) frac = __ddiv_rn((double)a[d], (double)c.total_resources[d]);
      v = __dmul_rn(frac, c.drf_mult[d]);
    }
    for (int off = 16; off > 0; off >>= 1) {
      double o = __shfl_xor_sync(FULL, v, off);
      v = o > v ? o : v;
    }
    return v > 0.0 ? v : 0.0;
  }

  // ---- alloc mutation on one node (lane = (level, resource) pair) ---------------------------
  // markAllocatable over levels <= cutoff with sign, plus optional EvictedPriority adjustment.
  __device__ void node_apply(uint32_t n, uint32_t cls, int32_t cutoff, int sign, int evicted_level_sign) {
    int idx = lane_id();
    for (; idx < c.PL * c.D; idx += 32) {
      int p = idx / c.D, d = idx % c.D;
      int64_t r = P.class_req_node[(size_t)cls * c.D + d];
      int64_t delta = 0;
      if (c.priorities[p] <= cutoff) delta += sign * r;
      if (p == 0) delta += evicted_level_sign * r;
      if (delta) P.alloc[((size_t)p * c.D + d) * c.N + n] += delta;
    }
    __threadfence_block();
    __syncwarp();
  }

  __device__ void undo_push(uint32_t kind, uint32_t job, uint32_t node, uint32_t extra) {
    if (lane_id() == 0) {
      uint32_t i = H->undo_n;
      P.undo_log[4 * (size_t)i + 0] = kind;
      P.undo_log[4 * (size_t)i + 1] = job;
      P.undo_log[4 * (size_t)i + 2] = node;
      P.undo_log[4 * (size_t)i + 3] = extra;
      H->undo_n = i + 1;
    }
    __syncwarp();
  }

  // bindJobToNodeInPlace (nodedb.go:915-949)
  __device__ void bind(uint32_t n, uint32_t job, int32_t priority) {
    uint32_t cls = P.job_class[job];
    bool is_evicted = P.bound_node[job] == n && P.evicted_on_node[job];
    int32_t cut = cutoff_for(c, P.class_pc[cls], priority);
    node_apply(n, cls, cut, -1, is_evicted ? +1 : 0);
    if (lane_id() == 0) {
      if (is_evicted) P.evicted_on_node[job] = 0;
      else P.bound_node[job] = n;
      P.sched_prio[job] = priority;  // not part of the memdb txn (nodedb.go:946)
    }
    undo_push(is_evicted ? 2u : 1u, job, n, (uint32_t)priority);
    uint32_t mask = levelmask_upto(cut);
    if (is_evicted) mask &= ~1u;  // EvictedPriority level nets to zero
    refresh(n, mask);
  }
  // unbindJobFromNodeInPlace for an evicted job (nodedb.go:1059-1098): only EvictedPriority moves
  __device__ void unbind_evicted(uint32_t job, uint32_t n) {
    uint32_t cls = P.job_class[job];
    node_apply(n, cls, INT32_MIN, 0, +1);
    if (lane_id() == 0) {
      P.evicted_on_node[job] = 0;
      P.bound_node[job] = NONE;
    }
    __syncwarp();
    refresh(n, 1u);
  }

  // ---- scheduling-context accounting (context/scheduling.go, context/queue.go) ----------------
  // qctx.addJobSchedulingContext + sctx.AddJobSchedulingContext; returns evictedInThisRound
  __device__ bool add_job(uint32_t job, bool successful, uint8_t reason) {
    uint32_t q = P.job_queue[job];
    uint32_t cls = P.job_class[job];
    uint32_t pc = P.class_pc[cls];
    bool ev = P.q_evicted[job];
    int d = lane_id();
    if (successful) {
      if (d < c.D) {
        int64_t r = P.class_req_raw[(size_t)cls * c.D + d];
        H->q_alloc[q][d] += r;
        P.q_alloc_pc[((size_t)q * c.PC + pc) * c.D + d] += r;
        if (ev) H->s_evicted[d] -= r;
        else H->s_scheduled[d] += r;
      }
      if (d == 0) {
        P.q_unsuccessful[job] = 0;
        if (ev) {
          P.q_evicted[job] = 0;
          P.q_rescheduled[job] = 1;
          H->num_evicted_jobs--;
        } else {
          P.q_successful[job] = 1;
          H->num_sched_jobs++;
        }
      }
    } else if (d == 0) {
      P.q_unsuccessful[job] = 1;
      P.q_reason[job] = reason;
    }
    __syncwarp();
    return ev;
  }
  // sctx.EvictJob + qctx.evictJob; returns scheduledInThisRound
  __device__ bool evict_job_ctx(uint32_t job) {
    uint32_t q = P.job_queue[job];
    uint32_t cls = P.job_class[job];
    uint32_t pc = P.class_pc[cls];
    bool sched = P.q_successful[job], resched = P.q_rescheduled[job];
    int d = lane_id();
    if (d < c.D) {
      int64_t r = P.class_req_raw[(size_t)cls * c.D + d];
      H->q_alloc[q][d] -= r;
      P.q_alloc_pc[((size_t)q * c.PC + pc) * c.D + d] -= r;
      if (sched) H->s_scheduled[d] -= r;
      else H->s_evicted[d] += r;
    }
    __syncwarp();
    if (d == 0) {
      if (sched || resched) {
        P.q_successful[job] = 0;
        P.q_rescheduled[job] = 0;
      } else {
        P.q_evicted[job] = 1;
      }
      if (sched) H->num_sched_jobs--;
      else H->num_evicted_jobs++;
    }
    __syncwarp();
    return sched;
  }

  // ---- gang member access ----------------------------------------------------------------------
  __device__ uint32_t gang_member(int q, uint32_t i) const {
    if (!H->ng_is_gang[q]) return H->ng_first[q];
    return P.gang_buf[P.gang_off[H->ng_first[q]] + i];
  }

  // ---- QueuedGangIterator (queue_scheduler.go:277-383) over MultiJobsIterator ------------------
  __device__ void it_only_yield_evicted(int q) {
    if (lane_id() == 0) {
      H->it_only_evicted[q] = 1;
      if (H->it_has_next[q] && !H->ng_all_evicted[q]) H->it_has_next[q] = 0;
    }
    __syncwarp();
  }
  // Peek: advance queue q's iterator to its next gang; returns false when exhausted.
  __device__ bool it_peek(int q, bool skip_known) {
    if (H->it_has_next[q]) return true;
    for (;;) {
      if (c.max_lookback != 0 && !H->it_only_evicted[q] && H->jobs_seen[q] >= c.max_lookback) it_only_yield_evicted(q);
      uint32_t job = NONE;
      bool from_evicted = false;
      if (H->ev_cur[q] < H->ev_end[q]) {
        job = P.evq_jobs[H->ev_cur[q]];
        from_evicted = true;
        if (lane_id() == 0) H->ev_cur[q]++;
        __syncwarp();
      } else if (a.with_queued && !H->it_only_evicted[q] && H->qd_cur[q] < H->qd_end[q]) {
        // warp-wide window over the queued jobs: skip known-unschedulable keys in bulk
        uint32_t cur = H->qd_cur[q], end = H->qd_end[q];
        uint32_t avail = end - cur;
        if (avail > 32) avail = 32;
        if (c.max_lookback != 0) {
          uint32_t room = c.max_lookback - H->jobs_seen[q];
          if (avail > room) avail = room;
        }
        bool can_skip = skip_known && H->num_unfeasible > 0;
        uint32_t jl = NONE;
        bool skippable = false;
        if ((uint32_t)lane_id() < avail) {
          jl = P.queued_order[cur + lane_id()];
          if (can_skip) {
            uint32_t cl = P.job_class[jl];
            skippable = P.class_key_valid[cl] && P.unfeasible[cl] != 0;
          }
        }
        uint32_t live = avail == 32 ? FULL : ((1u << avail) - 1u);
        uint32_t nonskip = __ballot_sync(FULL, (uint32_t)lane_id() < avail && !skippable) & live;
        uint32_t f = nonskip ? (uint32_t)(__ffs(nonskip) - 1) : avail;  // first job to process normally
        if ((uint32_t)lane_id() < f) {
          // jctx copies the failed jctx's reason; AddJobSchedulingContext ⇒ Unsuccessful
          P.q_unsuccessful[jl] = 1;
          P.q_reason[jl] = P.unfeasible[P.job_class[jl]];
          P.has_pctx[jl] = 1;
          P.res_node[jl] = NONE;
          P.res_method[jl] = 0;
        }
        uint32_t consumed = f < avail ? f + 1 : f;
        if (f < avail) job = __shfl_sync(FULL, jl, f);
        if (lane_id() == 0) {
          H->qd_cur[q] = cur + consumed;
          H->jobs_seen[q] += consumed;
        }
        __syncwarp();
        if (job == NONE) continue;  // whole window skipped
      } else {
        return false;
      }
      (void)from_evicted;
      // gang assembly
      uint32_t g = P.job_gang[job];
      if (g != NONE) {
        uint32_t fill = P.gang_fill[g];
        if (lane_id() == 0) {
          P.gang_buf[P.gang_off[g] + fill] = job;
          P.gang_fill[g] = fill + 1;
        }
        __syncwarp();
        if (fill + 1 == P.jc_card[job]) {
          if (lane_id() == 0) {
            P.gang_fill[g] = 0;
            H->ng_is_gang[q] = 1;
            H->ng_first[q] = g;
            H->ng_count[q] = fill + 1;
          }
          __syncwarp();
          finish_gang(q);
          return true;
        }
      } else {
        if (lane_id() == 0) {
          H->ng_is_gang[q] = 0;
          H->ng_first[q] = job;
          H->ng_count[q] = 1;
        }
        __syncwarp();
        finish_gang(q);
        return true;
      }
    }
  }
  // NewGangSchedulingContext (context/gang.go:22-49): totals + AllJobsEvicted
  __device__ void finish_gang(int q) {
    uint32_t cnt = H->ng_count[q];
    int d = lane_id();
    int64_t tot = 0;
    bool all_ev = true;
    for (uint32_t i = 0; i < cnt; ++i) {
      uint32_t j = gang_member(q, i);
      all_ev = all_ev && P.jc_evicted[j];
      if (d < c.D) tot += P.class_req_raw[(size_t)P.job_class[j] * c.D + d];
    }
    if (d < c.D) H->ng_total[q][d] = tot;
    if (d == 0) {
      H->ng_all_evicted[q] = all_ev;
      H->it_has_next[q] = 1;
    }
    __syncwarp();
  }

  // updatePQItem (queue_scheduler.go:536-580)
  __device__ void pq_update(int q, bool skip_known) {
    bool has = it_peek(q, skip_known);
    if (!has) {
      if (lane_id() == 0) H->q_in_pq[q] = 0;
      __syncwarp();
      return;
    }
    int d = lane_id();
    __shared__ int64_t tmp_base[ARMADA_MAX_RESOURCES], tmp_with[ARMADA_MAX_RESOURCES];
    if (d < c.D) {
      int64_t b = H->q_alloc[q][d] + P.queue_penalty[(size_t)q * c.D + d];
      tmp_base[d] = b;
      tmp_with[d] = b + H->ng_total[q][d];
    }
    __syncwarp();
    double w = P.queue_weight[q];
    double proposed = __ddiv_rn(drf_cost(tmp_with), w);
    double current = __ddiv_rn(drf_cost(tmp_base), w);
    double size = __dmul_rn(drf_cost(H->ng_total[q]), w);
    int32_t pr = INT32_MAX;
    uint32_t cnt = H->ng_count[q];
    for (uint32_t i = 0; i < cnt; ++i) {
      uint32_t j = gang_member(q, i);
      int32_t np = c.pcs[P.class_pc[P.job_class[j]]].priority;
      if (P.has_pctx[j]) np = P.res_sched_at[j];
      else if (P.job_node0[j] != NONE && P.job_sap0[j] != NOPRIO) np = P.job_sap0[j];
      if (np < pr) pr = np;
    }
    if (lane_id() == 0) {
      H->q_proposed[q] = proposed;
      H->q_current[q] = current;
      H->q_size[q] = size;
      H->q_pcprio[q] = pr;
      H->q_in_pq[q] = 1;
    }
    __syncwarp();
  }

  // QueueCandidateGangIteratorPQ.Less (queue_scheduler.go:628-674)
  __device__ bool item_less(const QItem& x, const QItem& y) const {
    if (y.q < 0) return x.q >= 0;
    if (x.q < 0) return false;
    if (a.consider_priority && x.pcprio != y.pcprio) return x.pcprio > y.pcprio;
    if (c.prefer_large) {
      bool xu = x.proposed <= x.budget, yu = y.proposed <= y.budget;
      if (xu && yu) {
        if (x.current == y.current && x.size != y.size) return x.size > y.size;
        if (x.current != y.current) return x.current < y.current;
      } else if (!xu && !yu) {
        if (x.proposed != y.proposed) return x.proposed < y.proposed;
      } else if (xu) {
        return true;
      } else if (yu) {
        return false;
      }
    } else {
      if (x.proposed != y.proposed) return x.proposed < y.proposed;
    }
    return x.q < y.q;
  }
  // heap top = arg-min over queues (warp reduction; `Less` is a strict total order)
  __device__ int pq_top() const {
    QItem best;
    best.q = -1;
    best.proposed = best.current = best.budget = best.size = 0;
    best.pcprio = 0;
    for (int q = lane_id(); q < c.Q; q += 32) {
      if (!H->q_in_pq[q]) continue;
      QItem it;
      it.q = q;
      it.proposed = H->q_proposed[q];
      it.current = H->q_current[q];
      it.budget = H->q_budget[q];
      it.size = H->q_size[q];
      it.pcprio = H->q_pcprio[q];
      if (item_less(it, best)) best = it;
    }
    for (int off = 16; off > 0; off >>= 1) {
      QItem o;
      o.proposed = __shfl_xor_sync(FULL, best.proposed, off);
      o.current = __shfl_xor_sync(FULL, best.current, off);
      o.budget = __shfl_xor_sync(FULL, best.budget, off);
      o.size = __shfl_xor_sync(FULL, best.size, off);
      o.pcprio = __shfl_xor_sync(FULL, best.pcprio, off);
      o.q = __shfl_xor_sync(FULL, best.q, off);
      if (item_less(o, best)) best = o;
    }
    return best.q;
  }

  // ---- constraints (constraints/constraints.go:113-178) ------------------------------------------
  __device__ uint8_t check_round_constraints() const {
    if (!c.has_round_limit) return 0;
    bool ex = false;
    for (int d = 0; d < c.D; ++d) ex = ex || H->s_scheduled[d] > c.max_to_schedule[d];
    return ex ? ARMADA_REASON_MAX_RESOURCES_SCHEDULED : 0;
  }
  __device__ uint8_t check_job_constraints(int q, uint32_t pc, uint32_t card) const {
    if (P.queue_cordoned[q]) return ARMADA_REASON_QUEUE_CORDONED;
    double tokens = H->global_tokens;
    if (tokens <= 0) return ARMADA_REASON_GLOBAL_RATE_LIMIT;
    if (c.global_burst < (int64_t)card) return ARMADA_REASON_GANG_EXCEEDS_GLOBAL_BURST;
    if (tokens < (double)card) return ARMADA_REASON_GLOBAL_RATE_LIMIT_GANG;
    tokens = H->q_tokens[q];
    if (tokens <= 0) return ARMADA_REASON_QUEUE_RATE_LIMIT;
    if (P.queue_burst[q] < (int64_t)card) return ARMADA_REASON_GANG_EXCEEDS_QUEUE_BURST;
    if (tokens < (double)card) return ARMADA_REASON_QUEUE_RATE_LIMIT_GANG;
    if (P.queue_has_limit[(size_t)q * c.PC + pc]) {
      bool ex = false;
      for (int d = 0; d < c.D; ++d)
        ex = ex || P.q_alloc_pc[((size_t)q * c.PC + pc) * c.D + d] > P.queue_limit[((size_t)q * c.PC + pc) * c.D + d];
      if (ex) return ARMADA_REASON_MAX_RESOURCES_PER_QUEUE;
    }
    return 0;
  }

  // ---- node selection ------------------------------------------------------------------------------
  // selectNodeForPodAtPriority (nodedb.go:717-805) on the aligned fast path: the first node of
  // the merged ordered walk that passes JobRequirementsMet == arg-min of the packed key.
  __device__ uint32_t probe(uint32_t cls, int variant, int level) {
    if (lane_id() == 0) H->st_probes++;
    uint32_t row = P.class_row[(size_t)cls * ARMADA_DEV_VARIANTS + variant];
    size_t si = ((size_t)cls * ARMADA_DEV_VARIANTS + variant) * c.PL + level;
    int s = P.slot_of[si];
    if (s == -1) {
      if (H->num_active_slots < c.num_slots) {
        s = H->num_active_slots;
        if (lane_id() == 0) {
          H->slot_level[s] = level;
          H->slot_row[s] = row;
          for (int d = 0; d < c.D; ++d) H->slot_req[s][d] = P.class_req_node[(size_t)cls * c.D + d];
          H->cmd_slot = s;
          H->num_active_slots = s + 1;
          P.slot_of[si] = (int16_t)s;
        }
        __threadfence_block();
        issue(CMD_BUILD);
      } else {
        s = -2;
        if (lane_id() == 0) P.slot_of[si] = -2;
        __syncwarp();
      }
    }
    unsigned long long key;
    if (s >= 0) {
      key = T.root[s];
    } else {
      if (lane_id() == 0) {
        H->cmd_level = level;
        H->cmd_row = row;
        for (int d = 0; d < c.D; ++d) H->cmd_req[d] = P.class_req_node[(size_t)cls * c.D + d];
      }
      __threadfence_block();
      issue(CMD_SCAN);
      key = H->cmd_result;
    }
    if (key == KEY_INF) return NONE;
    return (uint32_t)(key & ((1ull << c.node_bits) - 1ull));
  }

  // selectNodeForJobWithFairPreemption (nodedb.go:812-903): walk evicted jobs by descending index
  __device__ uint32_t select_fair_preemption(uint32_t job, uint32_t cls, int variant, int32_t* preempted_at) {
    if (lane_id() == 0) {
      H->st_fair++;
      H->fp_epoch++;
    }
    __syncwarp();
    uint32_t epoch = H->fp_epoch;
    uint32_t row = P.class_row[(size_t)cls * ARMADA_DEV_VARIANTS + variant];
    const int64_t* rq = P.class_req_node + (size_t)cls * c.D;
    for (int64_t i = (int64_t)a.num_evicted - 1; i >= 0; --i) {
      if (!P.ev_alive[i]) continue;
      uint32_t ej = P.ev_job_by_index[i];
      uint32_t n = P.bound_node[ej];
      if (n == NONE) {
        if (lane_id() == 0) H->error = 10;
        __syncwarp();
        return NONE;
      }
      bool fresh = P.fp_epoch[n] != epoch;
      int d = lane_id();
      if (!fresh && P.fp_bad[n]) continue;
      const int64_t* erq = P.class_req_node + (size_t)P.job_class[ej] * c.D;
      bool dyn_lane = true;
      if (d < c.D) {
        int64_t av = fresh ? P.alloc[((size_t)0 * c.D + d) * c.N + n] : P.fp_avail[(size_t)d * c.N + n];
        av += erq[d];
        P.fp_avail[(size_t)d * c.N + n] = av;
        dyn_lane = rq[d] <= av;
      }
      if (d == 0) {
        P.fp_next[i] = fresh ? NONE : P.fp_head[n];
        P.fp_head[n] = (uint32_t)i;
        if (fresh) {
          P.fp_epoch[n] = epoch;
          P.fp_bad[n] = 0;
        }
      }
      __threadfence_block();
      __syncwarp();
      bool dyn = __all_sync(FULL, dyn_lane);
      if (!dyn) continue;
      bool sok = static_ok_inline(c, P, row, rq, n);
      if (!sok) {
        if (d == 0) P.fp_bad[n] = 1;
        __syncwarp();
        continue;
      }
      // victims: every evicted job accumulated on this node
      int32_t maxp = -1;
      uint32_t cur = (uint32_t)i;
      // process in visiting order is not required (sums commute); walk the chain
      while (cur != NONE) {
        uint32_t pj = P.ev_job_by_index[cur];
        uint32_t nxt = P.fp_next[cur];
        unbind_evicted(pj, n);
        undo_push(3u, pj, n, cur);
        if (lane_id() == 0) {
          P.ev_alive[cur] = 0;
          P.ev_index_of_job[pj] = -1;
        }
        __syncwarp();
        int32_t pr = P.sched_prio[pj] != NOPRIO ? P.sched_prio[pj] : c.pcs[P.class_pc[P.job_class[pj]]].priority;
        if (pr > maxp) maxp = pr;
        cur = nxt;
      }
      *preempted_at = maxp;
      return n;
    }
    return NONE;
  }

  // selectNodeForJobWithTxnAtPriority (nodedb.go:605-666)
  __device__ uint32_t select_at_priority(uint32_t job, uint32_t cls, int variant, int32_t sched_at, int32_t* preempted_at,
                                          uint8_t* method) {
    uint32_t n = probe(cls, variant, 0);
    if (n != NONE) {
      *preempted_at = -1;
      *method = ARMADA_METHOD_NO_PREEMPTION;
      return n;
    }
    int lvl = level_of(sched_at);
    if (lvl < 0) {
      if (lane_id() == 0) H->error = 11;
      __syncwarp();
      return NONE;
    }
    n = probe(cls, variant, lvl);
    if (n == NONE) return NONE;
    n = select_fair_preemption(job, cls, variant, preempted_at);
    if (n != NONE) {
      *method = ARMADA_METHOD_FAIRSHARE;
      return n;
    }
    // selectNodeForJobWithUrgencyPreemption (nodedb.go:682-715)
    for (int p = 1; p < c.PL; ++p) {
      if (c.priorities[p] > sched_at) break;
      n = probe(cls, variant, p);
      if (n != NONE) {
        *preempted_at = c.priorities[p];
        *method = ARMADA_METHOD_URGENCY;
        return n;
      }
    }
    return NONE;
  }

  // SelectNodeForJobWithTxn (nodedb.go:431-512) + bind; returns false when the job does not fit
  __device__ bool select_and_bind(uint32_t job) {
    uint32_t cls = P.job_class[job];
    uint32_t pcidx = P.class_pc[cls];
    const ArmadaPriorityClass& pc = c.pcs[pcidx];
    int32_t priority = P.sched_prio[job] != NOPRIO ? P.sched_prio[job] : pc.priority;
    int32_t sched_at = priority, preempted_at = -1;
    uint8_t method = ARMADA_METHOD_NONE;
    uint32_t n = NONE;
    if (P.jc_evicted[job]) {  // pinned to its node; dynamic requirements only (:465-476, :774-783)
      uint32_t an = P.assigned_node[job];
      uint8_t fl = P.node_flags[an];
      bool ok;
      if ((fl & ARMADA_NODE_UNSCHEDULABLE) && (fl & ARMADA_NODE_OVERALLOCATED)) {
        ok = true;
      } else {
        int lvl = level_of(priority);
        bool okl = true;
        int d = lane_id();
        if (lvl < 0) okl = false;
        else if (d < c.D) okl = P.class_req_node[(size_t)cls * c.D + d] <= P.alloc[((size_t)lvl * c.D + d) * c.N + an];
        ok = __all_sync(FULL, okl);
      }
      method = ARMADA_METHOD_RESCHEDULED;
      if (ok) {
        n = an;
        preempted_at = priority;
      }
    } else {
      bool disallowed = false;
      for (int d = 0; d < c.D; ++d)
        if (((c.disallowed_mask >> d) & 1u) && P.class_req_raw[(size_t)cls * c.D + d] > 0) disallowed = true;
      if (!disallowed) {
        if (!c.disable_home) n = select_at_priority(job, cls, 0, sched_at, &preempted_at, &method);
        bool in_gang = P.job_gang[job] != NONE;
        bool away_disabled = c.disable_away || (in_gang && c.disable_gang_away);
        if (n == NONE && !away_disabled) {
          for (uint32_t k = 0; k < pc.num_away && n == NONE; ++k) {
            if (P.class_row[(size_t)cls * ARMADA_DEV_VARIANTS + 1 + k] == NONE) continue;
            sched_at = pc.away_priority[k];
            preempted_at = -1;
            n = select_at_priority(job, cls, 1 + (int)k, sched_at, &preempted_at, &method);
            if (n != NONE) method = ARMADA_METHOD_AWAY;
          }
        }
      }
    }
    if (lane_id() == 0) {
      P.has_pctx[job] = 1;
      P.res_node[job] = n;
      P.res_sched_at[job] = sched_at;
      P.res_preempted_at[job] = n != NONE ? preempted_at : -1;
      P.res_method[job] = method;
    }
    __syncwarp();
    if (n == NONE) return false;
    bind(n, job, sched_at);
    int32_t idx = P.ev_index_of_job[job];
    if (idx >= 0) {  // deleteEvictedJobSchedulingContextIfExistsWithTxn
      if (lane_id() == 0) {
        P.ev_alive[idx] = 0;
        P.ev_index_of_job[job] = -1;
      }
      undo_push(4u, job, 0, (uint32_t)idx);
    }
    if (lane_id() == 0) H->st_placements++;
    __syncwarp();
    return true;
  }

  // txn.Abort(): replay the undo log backwards
  __device__ void txn_abort() {
    for (uint32_t i = H->undo_n; i-- > 0;) {
      uint32_t kind = P.undo_log[4 * (size_t)i + 0], job = P.undo_log[4 * (size_t)i + 1];
      uint32_t n = P.undo_log[4 * (size_t)i + 2], extra = P.undo_log[4 * (size_t)i + 3];
      uint32_t cls = P.job_class[job];
      if (kind == 1u || kind == 2u) {
        int32_t cut = cutoff_for(c, P.class_pc[cls], (int32_t)extra);
        node_apply(n, cls, cut, +1, kind == 2u ? -1 : 0);
        if (lane_id() == 0) {
          if (kind == 2u) P.evicted_on_node[job] = 1;
          else P.bound_node[job] = NONE;
          if (H->st_placements) H->st_placements--;
        }
        __syncwarp();
        uint32_t mask = levelmask_upto(cut);
        if (kind == 2u) mask &= ~1u;
        refresh(n, mask);
      } else if (kind == 3u) {
        node_apply(n, cls, INT32_MIN, 0, -1);
        if (lane_id() == 0) {
          P.bound_node[job] = n;
          P.evicted_on_node[job] = 1;
          P.ev_alive[extra] = 1;
          P.ev_index_of_job[job] = (int32_t)extra;
        }
        __syncwarp();
        refresh(n, 1u);
      } else if (kind == 4u) {
        if (lane_id() == 0) {
          P.ev_alive[extra] = 1;
          P.ev_index_of_job[job] = (int32_t)extra;
        }
        __syncwarp();
      }
    }
    if (lane_id() == 0) H->undo_n = 0;
    __syncwarp();
  }

  // GangScheduler.Schedule (gang_scheduler.go:100-254) for queue q's current gang
  __device__ bool gang_schedule(int q, uint8_t* reason_out) {
    uint32_t cnt = H->ng_count[q];
    bool all_evicted = H->ng_all_evicted[q];
    uint8_t reason = 0;
    if (!all_evicted) {
      reason = check_round_constraints();
      if (reason) {  // returns before the deferred bookkeeping is registered (:102-109)
        *reason_out = reason;
        return false;
      }
    }
    // AddGangSchedulingContext (all jctx are "successful" until proven otherwise)
    bool all_ev_round = true;
    for (uint32_t i = 0; i < cnt; ++i) {
      uint32_t j = gang_member(q, i);
      if (lane_id() == 0) P.q_reason[j] = 0;
      bool ev = add_job(j, true, 0);
      all_ev_round = all_ev_round && ev;
    }
    if (!all_ev_round && lane_id() == 0) H->num_sched_gangs++;
    __syncwarp();
    bool ok = true;
    uint32_t pc = P.class_pc[P.job_class[gang_member(q, 0)]];
    if (!all_evicted) {
      reason = check_job_constraints(q, pc, cnt);
      if (reason) ok = false;
    }
    if (ok) {
      if (lane_id() == 0) H->undo_n = 0;
      __syncwarp();
      for (uint32_t i = 0; i < cnt && ok; ++i) ok = select_and_bind(gang_member(q, i));
      if (!ok) {
        txn_abort();
        reason = cnt > 1 ? ARMADA_REASON_GANG_DOES_NOT_FIT : ARMADA_REASON_JOB_DOES_NOT_FIT;
      }
      if (H->error) return false;
    }
    if (ok) {
      if (!all_evicted && lane_id() == 0) {  // Limiter.ReserveN (:118-123)
        if (!c.global_inf) H->global_tokens -= (double)cnt;
        if (!P.queue_inf[q]) H->q_tokens[q] -= (double)cnt;
      }
      __syncwarp();
    } else {
      // updateGangSchedulingContextOnFailure (:63-98)
      bool all_sched = true;
      for (uint32_t i = 0; i < cnt; ++i) all_sched = evict_job_ctx(gang_member(q, i)) && all_sched;
      if (all_sched && lane_id() == 0) H->num_sched_gangs--;
      __syncwarp();
      for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t j = gang_member(q, i);
        if (lane_id() == 0 && P.has_pctx[j]) {  // jctx.Fail
          P.res_node[j] = NONE;
          P.res_method[j] = 0;
        }
        add_job(j, false, reason);
      }
      bool prop = reason == ARMADA_REASON_GANG_EXCEEDS_GLOBAL_BURST || reason == ARMADA_REASON_JOB_DOES_NOT_FIT ||
                  reason == ARMADA_REASON_GANG_DOES_NOT_FIT;
      if (!a.skip_key_check && cnt == 1 && prop) {
        uint32_t j = gang_member(q, 0);
        uint32_t cl = P.job_class[j];
        if (!P.jc_evicted[j] && P.class_key_valid[cl] && !P.unfeasible[cl]) {
          if (lane_id() == 0) {
            P.unfeasible[cl] = reason;
            H->num_unfeasible++;
          }
          __threadfence_block();
          __syncwarp();
        }
      }
    }
    *reason_out = reason;
    return ok;
  }

  // Clear (queue_scheduler.go:495-506): pop the top item, advance its iterator, re-push
  __device__ void candidate_clear(int q, bool skip_known) {
    if (lane_id() == 0) {
      H->q_in_pq[q] = 0;
      H->it_has_next[q] = 0;
    }
    __syncwarp();
    pq_update(q, skip_known);
  }

  // addEvictedJobsToNodeDb (preempting_queue_scheduler.go:584-633): drain a cost-ordered gang
  // iterator over the evicted jobs with FROZEN allocations to assign the global index.
  __device__ void assign_evicted_indices() {
    int saved_cp = a.consider_priority;
    (void)saved_cp;
    // iterators over evicted lists only, no lookback, no key skipping; considerPriority=false
    reset_gang_fill();
    for (int q = lane_id(); q < c.Q; q += 32) {
      H->ev_cur[q] = P.evq_start[q];
      H->ev_end[q] = P.evq_start[q + 1];
      H->qd_cur[q] = H->qd_end[q] = 0;
      H->jobs_seen[q] = 0;
      H->it_only_evicted[q] = 1;  // disables the queued iterator and the lookback check
      H->it_has_next[q] = 0;
      H->q_in_pq[q] = 0;
    }
    __syncwarp();
    for (int q = 0; q < c.Q; ++q) pq_update(q, false);
    uint32_t idx = 0;
    for (;;) {
      int q = pq_top_no_priority();
      if (q < 0) break;
      uint32_t cnt = H->ng_count[q];
      for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t j = gang_member(q, i);
        if (lane_id() == 0) {
          P.ev_job_by_index[idx + i] = j;
          P.ev_index_of_job[j] = (int32_t)(idx + i);
          P.ev_alive[idx + i] = 1;
        }
      }
      idx += cnt;
      __syncwarp();
      candidate_clear(q, false);
    }
  }
  // every QueuedGangIterator starts with an empty jctxsByGangId map (queue_scheduler.go:294-301)
  __device__ void reset_gang_fill() {
    for (uint32_t g = lane_id(); g < c.G; g += 32) P.gang_fill[g] = 0;
    __threadfence_block();
    __syncwarp();
  }
  __device__ int pq_top_no_priority() {
    // same as pq_top with considerPriority=false
    PassArgs b = a;
    b.consider_priority = 0;
    Ctl t{c, P, H, T, b};
    return t.pq_top();
  }
};

__global__ void __launch_bounds__(1024, 1) k_schedule_pass(DevCfg c, DevPtrs P, PassArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemHdr* H = (SmemHdr*)smem_raw;
  Trees T;
  size_t off = (sizeof(SmemHdr) + 15) & ~(size_t)15;
  T.root = (unsigned long long*)(smem_raw + off);
  off += sizeof(unsigned long long) * ARMADA_DEV_MAX_SLOTS;
  T.l2 = (unsigned long long*)(smem_raw + off);
  off += sizeof(unsigned long long) * (size_t)c.num_slots * c.num_groups;
  T.leaf = (unsigned long long*)(smem_raw + off);

  if (threadIdx.x == 0) {
    H->cmd = CMD_EXIT;
    H->num_active_slots = 0;
    H->num_unfeasible = 0;
    H->only_yield_evicted = 0;
    H->undo_n = 0;
    H->fp_epoch = 0;
    H->error = 0;
    H->st_iters = H->st_probes = H->st_placements = H->st_fair = H->st_rescans = 0;
    H->num_sched_jobs = P.s_counts[0];
    H->num_sched_gangs = P.s_counts[1];
    H->num_evicted_jobs = P.s_counts[2];
    H->global_tokens = __longlong_as_double(P.s_counts[4]);
    for (int d = 0; d < c.D; ++d) {
      H->s_scheduled[d] = P.s_scheduled[d];
      H->s_evicted[d] = P.s_evicted[d];
    }
  }
  for (int q = threadIdx.x; q < c.Q; q += blockDim.x) {
    for (int d = 0; d < c.D; ++d) H->q_alloc[q][d] = P.q_alloc[(size_t)q * c.D + d];
    H->q_tokens[q] = P.q_tokens[q];
    H->q_budget[q] = __ddiv_rn(P.q_fair[q * 3 + 1], P.queue_weight[q]);  // queue_scheduler.go:438
    H->q_in_pq[q] = 0;
    H->it_has_next[q] = 0;
    H->q_only_evicted_for_queue[q] = 0;
  }
  // trees are rebuilt lazily per pass: forget the slot table
  for (size_t i = threadIdx.x; i < (size_t)c.C * ARMADA_DEV_VARIANTS * c.PL; i += blockDim.x) P.slot_of[i] = -1;
  for (size_t i = threadIdx.x; i < (size_t)c.C; i += blockDim.x) P.unfeasible[i] = 0;  // ClearUnfeasibleSchedulingKeys
  __threadfence_block();
  __syncthreads();

  if (warp_id() != 0) {  // worker warps
    for (;;) {
      bar_all();
      if (H->cmd == CMD_EXIT) break;
      do_work(c, P, H, T);
      bar_all();
    }
    return;
  }

  Ctl ctl{c, P, H, T, a};
  if (a.assign_indices) ctl.assign_evicted_indices();

  // NewQueueScheduler: per-queue MultiJobsIterator(evicted, queued) + cost PQ (:66-90, :404-443)
  ctl.reset_gang_fill();
  for (int q = lane_id(); q < c.Q; q += 32) {
    H->ev_cur[q] = P.evq_start[q];
    H->ev_end[q] = P.evq_start[q + 1];
    H->qd_cur[q] = a.with_queued ? P.queued_start[q] : 0;
    H->qd_end[q] = a.with_queued ? P.queued_start[q + 1] : 0;
    H->jobs_seen[q] = 0;
    H->it_only_evicted[q] = 0;
    H->it_has_next[q] = 0;
    H->q_in_pq[q] = 0;
  }
  __syncwarp();
  for (int q = 0; q < c.Q; ++q) ctl.pq_update(q, true);

  uint32_t term = 0;
  for (;;) {  // QueueScheduler.Schedule loop (queue_scheduler.go:100-236)
    if (H->error) break;
    int q = ctl.pq_top();
    if (q < 0) break;
    if (lane_id() == 0) H->st_iters++;
    uint8_t reason = 0;
    bool ok = ctl.gang_schedule(q, &reason);
    if (H->error) break;
    uint32_t cnt = H->ng_count[q];
    // remember the gang for the result before Clear() advances the iterator
    if (ok) {
      for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t j = ctl.gang_member(q, i);
        if (lane_id() == 0 && P.res_node[j] != NONE) {
          // result algebra of preempting_queue_scheduler.go:157-163 / :208-215
          if (P.in_preempted[j]) P.in_preempted[j] = 0;
          else P.in_scheduled[j] = 1;
          if (a.pass == 2) P.in_sae[j] = 0;
        }
      }
      __syncwarp();
    }
    ctl.candidate_clear(q, true);
    if (!ok) {
      if (reason == ARMADA_REASON_MAX_RESOURCES_SCHEDULED || reason == ARMADA_REASON_GLOBAL_RATE_LIMIT) {
        term = reason;  // IsTerminalUnschedulableReason ⇒ OnlyYieldEvicted (:446-469)
        if (!H->only_yield_evicted) {
          for (int qq = 0; qq < c.Q; ++qq) {
            if (!H->q_in_pq[qq]) continue;
            ctl.it_only_yield_evicted(qq);
            ctl.pq_update(qq, true);
          }
          if (lane_id() == 0) H->only_yield_evicted = 1;
          __syncwarp();
        }
      } else if (reason == ARMADA_REASON_QUEUE_RATE_LIMIT || reason == ARMADA_REASON_QUEUE_CORDONED) {
        // OnlyYieldEvictedForQueue (:471-491)
        if (!H->only_yield_evicted && !H->q_only_evicted_for_queue[q] && H->q_in_pq[q]) {
          ctl.it_only_yield_evicted(q);
          ctl.pq_update(q, true);
        }
        if (lane_id() == 0) H->q_only_evicted_for_queue[q] = 1;
        __syncwarp();
      }
    }
  }
  if (term == 0) term = ARMADA_REASON_NO_REMAINING_CANDIDATES;

  // publish state for the next kernel
  for (int q = lane_id(); q < c.Q; q += 32) {
    for (int d = 0; d < c.D; ++d) P.q_alloc[(size_t)q * c.D + d] = H->q_alloc[q][d];
    P.q_tokens[q] = H->q_tokens[q];
  }
  if (lane_id() == 0) {
    for (int d = 0; d < c.D; ++d) {
      P.s_scheduled[d] = H->s_scheduled[d];
      P.s_evicted[d] = H->s_evicted[d];
    }
    P.s_counts[0] = H->num_sched_jobs;
    P.s_counts[1] = H->num_sched_gangs;
    P.s_counts[2] = H->num_evicted_jobs;
    if (a.pass == 1) P.s_counts[3] = term;
    P.s_counts[4] = __double_as_longlong(H->global_tokens);
    P.s_counts[5] = H->error;
    P.stats[0] += H->st_iters;
    P.stats[1] += H->st_probes;
    P.stats[2] += H->st_placements;
    P.stats[3] += H->st_fair;
    P.stats[4] += H->st_rescans;
    H->cmd = CMD_EXIT;
  }
  __syncwarp();
  bar_all();
}

// unbindJobs(preempted ∪ scheduledAndEvicted) (preempting_queue_scheduler.go:256, :766-789) and
// the per-job outcome.
__global__ void k_finalize(DevCfg c, DevPtrs P, uint8_t* out_state, uint32_t* out_node, int32_t* out_sched_at,
                           int32_t* out_preempted_at, uint8_t* out_method, uint8_t* out_reason) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= c.J) return;
  uint8_t st = ARMADA_JOB_NONE;
  uint32_t node = NONE;
  bool pre = P.in_preempted[j], sae = P.in_sae[j];
  if (pre || sae) {
    uint32_t n = P.assigned_node[j];
    if (P.bound_node[j] == n && n != NONE) {  // unbindJobFromNodeInPlace
      uint32_t cls = P.job_class[j];
      bool is_ev = P.evicted_on_node[j];
      int32_t cut = cutoff_for(c, P.class_pc[cls], P.sched_prio[j]);
      for (int p = 0; p < c.PL; ++p) {
        bool hit = is_ev ? (p == 0) : (c.priorities[p] <= cut);
        if (!hit) continue;
        for (int d = 0; d < c.D; ++d) {
          int64_t r = P.class_req_node[(size_t)cls * c.D + d];
          if (r) atomic_add_i64(&P.alloc[((size_t)p * c.D + d) * c.N + n], r);
        }
      }
      P.bound_node[j] = NONE;
      P.evicted_on_node[j] = 0;
    }
  }
  if (pre) {
    st = ARMADA_JOB_PREEMPTED;
    node = P.assigned_node[j];
  } else if (P.in_scheduled[j]) {
    st = ARMADA_JOB_SCHEDULED;
    node = P.res_node[j];
  } else if (sae) {
    st = ARMADA_JOB_SCHEDULED_AND_EVICTED;
    node = P.assigned_node[j];
  } else if (P.q_rescheduled[j]) {
    st = ARMADA_JOB_RESCHEDULED;
    node = P.res_node[j];
  } else if (P.q_unsuccessful[j]) {
    st = ARMADA_JOB_FAILED;
  }
  bool has = P.has_pctx[j] && st != ARMADA_JOB_NONE && st != ARMADA_JOB_FAILED;
  out_state[j] = st;
  out_node[j] = node;
  out_sched_at[j] = has ? P.res_sched_at[j] : NOPRIO;
  out_preempted_at[j] = has ? P.res_preempted_at[j] : NOPRIO;
  out_method[j] = has ? P.res_method[j] : 0;
  out_reason[j] = P.q_unsuccessful[j] ? P.q_reason[j] : 0;
}

}  // namespace

#include "armada_host.inc"
