"""Rounds that exercise GangScheduler.trySchedule's node-uniformity search and the floating-resource limit
(gang_scheduler.go:143,154-223): the reference's TestGangScheduler table (tests/golden/gang_scheduler.json) as
rounds, and a seeded generator of larger ones.  Shared by the oracle, emulator and GPU test modules."""
from __future__ import annotations

import json
import random

import pytest

import fixtures as fx
import go_tables as gt
from armada_b200 import abi
from armada_b200.model import FloatingResource, QueueSpec, RoundInputBuilder

GANG = gt.load_cases("gang_scheduler")


def _away_node_types(F):
    """"AwayNodeTypes" (gang_scheduler_test.go:415-483): config, node and jobs are built in closures there."""
    from armada_b200.model import AwayNodeType, PriorityClass, Taint
    ta, tb = Taint("taint-a", "true", "NoSchedule"), Taint("taint-b", "true", "NoSchedule")
    cfg = fx.test_scheduling_config(
        priority_classes={"armada-preemptible-away": PriorityClass(30000, True, (AwayNodeType(29000, "node-type-a"), AwayNodeType(29000, "node-type-b"))),
                          "armada-preemptible-away-both": PriorityClass(30000, True, (AwayNodeType(29000, "node-type-ab"),))},
        well_known_node_types={"node-type-a": (ta,), "node-type-b": (tb,), "node-type-ab": (ta, tb)})
    node = F.gpu8()
    node.taints = node.taints + (ta, tb)
    jobs = [F.n_1cpu_4gi("A", "armada-preemptible-away", 1), F.n_1cpu_4gi("A", "armada-preemptible-away-both", 1)]
    return {"SchedulingConfig": cfg, "Nodes": [node], "Gangs": jobs, "ExpectedScheduledIndices": [1], "ExpectedCumulativeScheduledJobs": [0, 1]}


def _home_away(F):
    """"Home-away scheduling" (gang_scheduler_test.go:484-529)."""
    from armada_b200.model import AwayNodeType, PriorityClass, Taint
    ta = Taint("taint-a", "true", "NoSchedule")
    cfg = fx.test_scheduling_config(
        priority_classes={"armada-preemptible": PriorityClass(30000, True),
                          "armada-preemptible-away": PriorityClass(30000, True, (AwayNodeType(29000, "node-type-a"),))},
        well_known_node_types={"node-type-a": (ta,)})
    node = F.cpu32()
    node.taints = node.taints + (ta,)
    jobs = [F.n_32cpu_256gi_large_toleration("A", "armada-preemptible-away", 1), F.n_32cpu_256gi_large_toleration("A", "armada-preemptible-away", 1)]
    return {"SchedulingConfig": cfg, "Nodes": [node], "Gangs": jobs, "ExpectedScheduledIndices": [0], "ExpectedCumulativeScheduledJobs": [1, 1]}


def _queue_rate_limit(F):
    """"Hitting queue constraint does not make job scheduling key unfeasible" (gang_scheduler_test.go:530-547): the per-queue
    burst of 2 lets queue A's first gang through and stops its second; queue B is not affected.  (Two queues: whatever the
    order the round visits them in, the scheduled set is the same.)"""
    cfg = fx.test_scheduling_config(maximum_per_queue_scheduling_burst=2, maximum_per_queue_scheduling_rate=2.0)
    jobs = [fx.with_gang(F.n_1cpu_4gi("A", fx.PriorityClass0, 2)), F.n_1cpu_4gi("A", fx.PriorityClass0, 1), F.n_1cpu_4gi("B", fx.PriorityClass0, 1)]
    return {"SchedulingConfig": cfg, "Nodes": F.n_cpu32(1), "Gangs": jobs, "ExpectedScheduledIndices": [0, 2], "ExpectedCumulativeScheduledJobs": [2, 2, 3],
            "AnyQueueOrder": True}


MANUAL = {"AwayNodeTypes": _away_node_types, "Home-away scheduling": _home_away,
          "Hitting queue constraint does not make job scheduling key unfeasible": _queue_rate_limit}


def gang_case_round(name):
    """(builder, evaluated case, gangs) of one TestGangScheduler case; pytest.skip for what a round cannot express."""
    env = gt.Env()

    def with_uniformity(jobs, label):  # testfixtures.WithNodeUniformityGangAnnotationsJobs
        jobs = fx.with_gang(jobs)
        for j in jobs:
            j.gang_node_uniformity_label = label
        return jobs

    def add_floating(request, jobs):  # addFloatingResourceRequest (gang_scheduler_test.go:734-742)
        for j in jobs:
            j.requests = dict(j.requests)
            j.requests["test-floating-resource"] = request
        return jobs

    env.calls["testfixtures.WithNodeUniformityGangAnnotationsJobs"] = with_uniformity
    env.calls["addFloatingResourceRequest"] = add_floating
    src = GANG[name]
    if name in MANUAL:
        tc = MANUAL[name](env.fx)
    elif name == "floating resources":
        # the table builds this config in a closure (gang_scheduler_test.go:104-108): TestSchedulingConfig with
        # FloatingResources = TestFloatingResourceConfig (testfixtures.go:65-75: 10 of test-floating-resource in "pool")
        src = json.loads(json.dumps(src))
        src["elems"] = [e for e in src["elems"] if e[0].get("id") != "SchedulingConfig"]
        tc = env.ev(src)
        tc["SchedulingConfig"] = fx.test_scheduling_config(floating_resources=[FloatingResource("test-floating-resource", "1", "10")])
    else:
        try:
            tc = env.ev(src)
        except gt.UnsupportedCase as e:
            pytest.skip(f"not modelled: {e}")
    gangs = tc["Gangs"]
    if len({j.queue for g in gangs for j in g}) != 1 and not tc.get("AnyQueueOrder"):
        pytest.skip("multi-queue gang case: direct GangScheduler order differs from queue order")
    if tc.get("AddAwayQueueContexts"):
        pytest.skip("away queue contexts")
    cfg = tc["SchedulingConfig"]
    nodes = tc["Nodes"]
    t = 0
    jobs = []
    for g in gangs:
        for j in g:
            t += 1
            j.submit_time = t
            if len(g) == 1:
                j.gang_id, j.gang_cardinality = None, 1
            jobs.append(j)
    try:
        synth = gt.materialize_used(cfg, nodes, env.fx)
    except gt.UnsupportedCase as e:
        pytest.skip(str(e))
    b = RoundInputBuilder(cfg, nodes, jobs + synth, [QueueSpec(q, 1.0) for q in sorted({j.queue for j in jobs})])
    return b, tc, gangs


def check_gang_case(b, tc, gangs, res):
    got = []
    for gi, g in enumerate(gangs):
        st = [int(res.job_state[b.job_pos[j.id]]) for j in g]
        if all(s == abi.JOB_SCHEDULED for s in st):
            got.append(gi)
        else:
            assert all(s != abi.JOB_SCHEDULED for s in st), "gang partially scheduled"
    assert got == sorted(tc.get("ExpectedScheduledIndices") or [])
    for gi, value in (tc.get("ExpectedNodeUniformity") or {}).items():  # the value the uniformity search settled on
        label = gangs[gi][0].gang_node_uniformity_label
        assert {b.nodes[int(res.job_node[b.job_pos[j.id]])].labels.get(label) for j in gangs[gi]} == {value}
    cum = tc.get("ExpectedCumulativeScheduledJobs")
    if cum:
        assert int(res.out.num_scheduled_jobs) == int(cum[-1])


def uniformity_round(seed: int, n_nodes: int = 40, n_zones: int = 5, n_queues: int = 4, n_jobs: int = 260, floating: bool = True,
                     unaligned: bool = False, max_gang: int = 6, running_gangs: int = 0, protected_fraction: float = 0.5) -> RoundInputBuilder:
    """Nodes spread over `n_zones` values of the label "zone" (a few nodes without it), partly filled with running
    preemptible jobs of an over-served queue (so attempts differ in what they preempt: the fit's mean
    PreemptedAtPriority), queued single jobs and gangs — some gangs with the uniformity label "zone", some with a label
    that is not indexed, some with a label no node carries; optionally a floating resource "licences" with a pool limit
    that the queued jobs exhaust."""
    rnd = random.Random(seed)
    F = fx.Fixtures()
    cfg = fx.test_scheduling_config(indexed_node_labels=list(fx.test_scheduling_config().indexed_node_labels) + ["zone", "rack"],
                                    protected_fraction_of_fair_share=protected_fraction)
    if floating:
        cfg.floating_resources = [FloatingResource("licences", "1", str(rnd.randint(6, 30)))]
    nodes = []
    for i in range(n_nodes):
        labels = {} if rnd.random() < 0.1 else {"zone": f"z{rnd.randrange(n_zones)}"}
        nodes.append(F.node({"cpu": "32", "memory": "256Gi"}, labels=labels))
    mem = "4100Mi" if unaligned else "4Gi"  # 4100Mi is not a multiple of the 128Mi index resolution: exact mode
    jobs = []
    # running jobs of queue "a" (preemptible priorities 0..2) fill most nodes
    for n in nodes:
        for _ in range(rnd.randint(0, 3)):
            pc = rnd.choice([fx.PriorityClass0, fx.PriorityClass1, fx.PriorityClass2])
            j = F.job("a", pc, {"cpu": str(rnd.choice([4, 8, 12])), "memory": "32Gi"})
            j.node = n.id
            j.scheduled_at_priority = fx.test_priority_classes()[pc].priority
            j.active_run_timestamp = len(jobs)
            jobs.append(j)
    queues = ["a"] + [f"q{i}" for i in range(n_queues - 1)]
    gid = 0
    # running gangs with a uniformity label (evicted and re-scheduled as gangs: every member is pinned to its node)
    for g in range(running_gangs):
        zone_nodes = [n for n in nodes if n.labels.get("zone") == f"z{g % n_zones}"]
        if len(zone_nodes) < 2:
            continue
        pc = rnd.choice([fx.PriorityClass0, fx.PriorityClass1])
        card = rnd.randint(2, 4)
        gid += 1
        for m in range(card):
            j = F.job("a", pc, {"cpu": "2", "memory": "8Gi"})
            j.node = zone_nodes[m % len(zone_nodes)].id
            j.scheduled_at_priority = fx.test_priority_classes()[pc].priority
            j.active_run_timestamp = len(jobs)
            j.gang_id, j.gang_cardinality, j.gang_node_uniformity_label = f"rg{gid}", card, "zone"
            jobs.append(j)
    while len(jobs) < n_jobs:
        q = rnd.choice(queues[1:])
        pc = rnd.choice([fx.PriorityClass0, fx.PriorityClass1, fx.PriorityClass2, fx.PriorityClass3])
        kind = rnd.random()
        req = {"cpu": str(rnd.choice([1, 2, 8, 16])), "memory": mem}
        if floating and rnd.random() < 0.3:
            req["licences"] = str(rnd.randint(1, 4))
        if kind < 0.45:
            jobs.append(F.job(q, pc, req))
            continue
        card = rnd.randint(2, max_gang)
        gid += 1
        label = None
        r = rnd.random()
        if r < 0.6:
            label = "zone"
        elif r < 0.7:
            label = "not-indexed-label"
        elif r < 0.8:
            label = "rack"  # indexed, on no node
        members = [F.job(q, pc, dict(req)) for _ in range(card)]
        for m in members:
            m.gang_id, m.gang_cardinality, m.gang_node_uniformity_label = f"g{gid}", card, label
        jobs += members
    # the snapshot's queue accounting (calculateJobSchedulingInfo): allocation = the running jobs, demand = every job
    import numpy as np
    f = cfg.factory()
    qs = []
    for q in queues:
        alloc, demand = {}, np.zeros(f.D, np.int64)
        for j in jobs:
            if j.queue != q:
                continue
            r = f.from_job(j.requests)
            demand += r
            if j.node is not None:
                alloc[j.priority_class] = alloc.get(j.priority_class, np.zeros(f.D, np.int64)) + r
        qs.append(QueueSpec(q, rnd.choice([1.0, 2.0]), False, alloc, demand.copy(), demand.copy()))
    return RoundInputBuilder(cfg, nodes, jobs, qs)
