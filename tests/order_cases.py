"""TestJobPriorityComparer (jobdb/comparison_test.go:13-75) as rounds: two jobs of one queue on a node that holds
both; the loop iteration each was attempted in (`job_seq`) is the order SchedulingOrderCompare put them in.  Running
jobs take part as evicted jobs (ProtectedFractionOfFairShare = 0 evicts every preemptible job; the evicted list is
ranked by the same comparator, jobiteration.go:72-95)."""
from __future__ import annotations

import numpy as np

import fixtures as fx
from armada_b200.model import JobSpec, QueueSpec, RoundInputBuilder

# name -> (a, b, expected Compare(a, b)); a job = dict(pc priority, queue priority, submit, run timestamp or None)
CASES = {
    "Queued jobs are ordered first by increasing priority class priority": (dict(pc=1, qp=1), dict(pc=2, qp=2), 1),
    "Queued jobs are ordered second by decreasing priority": (dict(pc=1, qp=2), dict(pc=1, qp=1), 1),
    "Queued jobs are ordered third by decreasing submit time": (dict(pc=1, qp=1, sub=2), dict(pc=1, qp=1, sub=1), 1),
    "Queued jobs are not ordered by runtime": (dict(pc=1, qp=1, sub=2, art=1), dict(pc=1, qp=1, sub=1, art=2), 1),
    "Queued jobs are ordered fourth by increasing id": (dict(pc=1, qp=1, sub=1), dict(pc=1, qp=1, sub=1), -1),
    "Running jobs come before queued jobs": (dict(pc=0, qp=1), dict(pc=0, qp=2, run=0), 1),
    "Running jobs are ordered third by runtime": (dict(pc=1, qp=1, sub=1, run=1), dict(pc=1, qp=1, sub=2, run=0), 1),
}
PC_OF = {0: fx.PriorityClass0, 1: fx.PriorityClass1, 2: fx.PriorityClass2}


def comparison_round(name):
    a, b, expected = CASES[name]
    F = fx.Fixtures()
    cfg = fx.test_scheduling_config(protected_fraction_of_fair_share=0.0)
    node = F.cpu32()
    jobs = []
    for jid, d in (("a", a), ("b", b)):
        pc = PC_OF[d["pc"]]
        j = JobSpec(id=jid, queue="A", priority_class=pc, requests={"cpu": "1", "memory": "4Gi"}, queue_priority=d.get("qp", 0),
                    submit_time=d.get("sub", 0), active_run_timestamp=d.get("art", 0))
        if "run" in d:  # WithNewRun / WithUpdatedRun: an active run
            j.node = node.id
            j.scheduled_at_priority = fx.test_priority_classes()[pc].priority
            j.active_run_timestamp = d["run"]
        jobs.append(j)
    # a second queue with demand and no allocation, so that queue A (everything it runs) is above its protected share
    other = F.job("B", fx.PriorityClass0, {"cpu": "1", "memory": "4Gi"})
    # the queue accounting of the snapshot (calculateJobSchedulingInfo): allocation = the running jobs, demand = all jobs
    f = cfg.factory()
    alloc, demand = {}, np.zeros(f.D, np.int64)
    for j in jobs:
        r = f.from_job(j.requests)
        demand += r
        if j.node is not None:
            alloc[j.priority_class] = alloc.get(j.priority_class, np.zeros(f.D, np.int64)) + r
    qa = QueueSpec("A", 1.0, False, alloc, demand.copy(), demand.copy())
    db = f.from_job(other.requests)
    b_ = RoundInputBuilder(cfg, [node], jobs + [other], [qa, QueueSpec("B", 1.0, False, {}, db.copy(), db.copy())])
    return b_, expected


def check_order(b, expected, res):
    sa, sb = int(res.job_seq[b.job_pos["a"]]), int(res.job_seq[b.job_pos["b"]])
    assert sa > 0 and sb > 0, "both jobs must have been attempted"
    assert (sa < sb) == (expected < 0), f"attempt order {sa} vs {sb}, Compare(a, b) = {expected}"
