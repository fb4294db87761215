"""Hand transcription of the reference's home / away node scheduling tests
(internal/scheduler/nodedb/nodedb_test.go): TestHomeNodeScheduling (:722-821),
TestConditionalAwayNodeScheduling (:908-964), TestAwayNodeScheduling (:966-1108) — replayed through
the oracle's NodeDb (ScheduleManyWithTxn) and, as one-gang rounds, through the CUDA source under the
SIMT emulator.  Every case cites the reference line of its expectation."""
import numpy as np
import pytest

import emu_lib
import fixtures as fx
import oracle_lib
from armada_b200 import abi
from armada_b200.model import QueueSpec, RoundInputBuilder, Taint, Toleration

GPU = Taint("gpu", "true", "NoSchedule")
LARGE = Taint("largeJobsOnly", "true", "NoSchedule")


def _case(cfg, node, jobs):
    b = RoundInputBuilder(cfg, [node], jobs, [QueueSpec(jobs[0].queue, 1.0)])
    db = oracle_lib.OracleNodeDb(b.input)
    ok, nodes, sched_at, _, method = db.schedule_many([b.job_pos[j.id] for j in jobs])
    return b, ok, nodes, sched_at, method


def _round_agrees(b, ok, method, sched_at):
    """The same gang as a scheduling round: oracle and emulated kernel agree, and the round places the
    gang exactly when the NodeDb call did (same method / priority)."""
    want = oracle_lib.round_schedule(b.input)
    got = emu_lib.emu_round().schedule(b.input)
    assert not got.diff(want)
    st = np.asarray(want.job_state)
    assert bool((st == abi.JOB_SCHEDULED).all()) == ok
    if ok:
        assert (np.asarray(want.job_method) == method[0]).all()
        assert (np.asarray(want.job_scheduled_at_priority) == sched_at[0]).all()


HOME = {  # nodedb_test.go:731-752
    "should schedule home jobs": dict(ok=True),
    "should schedule home jobs - when job toleration matches node taint": dict(ok=True, taint=True, toleration=True),
    "should not schedule home jobs - when job toleration does not match node taint": dict(ok=False, taint=True),
    "should not schedule home job - when home scheduling disabled": dict(ok=False, no_home=True),
    "should schedule away job - when home scheduling disabled - away scheduling available":
        dict(ok=True, away=True, no_home=True, pc=fx.PriorityClass4PreemptibleAway),
}


@pytest.mark.parametrize("name", sorted(HOME))
def test_home_node_scheduling(name):
    tc = HOME[name]
    f = fx.Fixtures()
    cfg = fx.test_scheduling_config(disable_home_scheduling=tc.get("no_home", False))
    cfg.well_known_node_types = {"gpu": (GPU,), "large": (LARGE,)}  # :762-765
    node = f.cpu32()  # Test32CpuNode([29000, 30000]) :775
    if tc.get("taint"):
        node.taints = node.taints + (LARGE,)  # :776-781
    pc = tc.get("pc", fx.PriorityClass6Preemptible)
    tol = (Toleration("largeJobsOnly", "", "true"),) if tc.get("toleration") else ()
    job = f.job("A", pc, {"cpu": "16", "memory": "128Gi"}, tol)  # Test16Cpu128GiJob[WithLargeJobToleration] :788-791
    b, ok, nodes, sched_at, method = _case(cfg, node, [job])
    assert ok == tc["ok"]  # :800 / :812
    if ok:
        assert nodes[0] == 0  # :804
        if tc.get("away"):
            assert method[0] == abi.METHOD_AWAY and sched_at[0] == 29000  # :806-807
        else:
            assert method[0] == abi.METHOD_NO_PREEMPTION and sched_at[0] == 30000  # :809-810
    _round_agrees(b, ok, method, sched_at)


def _gpu8(f, taint):
    n = f.gpu8()  # Test8GpuNode(TestPriorities)
    n.taints = n.taints + (taint,)
    return n


COND = {  # nodedb_test.go:920-943
    "scheduled - cpu job can conditionally schedule away on gpu nodes": dict(gpu_job=False, taint=GPU, ok=True),
    "not scheduled - gpu job can not conditionally schedule away on gpu nodes": dict(gpu_job=True, taint=GPU, ok=False),
    "scheduled - cpu job schedule away on large nodes": dict(gpu_job=False, taint=LARGE, ok=True),
    "scheduled - gpu job schedule away on large nodes": dict(gpu_job=True, taint=LARGE, ok=True),
}


@pytest.mark.parametrize("name", sorted(COND))
def test_conditional_away_node_scheduling(name):
    tc = COND[name]
    f = fx.Fixtures()
    cfg = fx.test_scheduling_config()
    node = _gpu8(f, tc["taint"])
    pc = fx.PriorityClass7PreemptibleAwayConditional
    req = {"cpu": "8", "memory": "128Gi", "nvidia.com/gpu": "1"} if tc["gpu_job"] else {"cpu": "1", "memory": "4Gi"}  # :909-914
    job = f.job("A", pc, req)  # (no gpu toleration)
    b, ok, nodes, sched_at, method = _case(cfg, node, [job])
    assert ok == tc["ok"]  # :955 / :960
    if ok:
        assert nodes[0] == 0 and method[0] == abi.METHOD_AWAY  # :956-958
    _round_agrees(b, ok, method, sched_at)


AWAY = {  # nodedb_test.go:975-1031
    "should schedule away jobs": dict(ok=True, node=GPU, wk=GPU),
    "should schedule away jobs - wild card well known node type": dict(ok=True, node=GPU, wk=Taint("gpu", "*", "NoSchedule")),
    "should schedule away jobs - gang": dict(gang=True, ok=True, node=GPU, wk=GPU),
    "should not schedule away jobs - when node doesn't match configured away types": dict(ok=False, node=Taint("cpu", "true", "NoSchedule"), wk=GPU),
    "should not schedule away jobs - when away scheduling disabled": dict(no_away=True, ok=False, node=GPU, wk=GPU),
    "should not schedule away jobs - gang - when away scheduling disabled": dict(gang=True, no_away=True, ok=False, node=GPU, wk=GPU),
    "should schedule away jobs - when gang away scheduling disabled": dict(no_gang_away=True, ok=True, node=GPU, wk=GPU),
    "should not schedule away jobs - gang - when gang away scheduling disabled": dict(gang=True, no_gang_away=True, ok=False, node=GPU, wk=GPU),
}


@pytest.mark.parametrize("name", sorted(AWAY))
def test_away_node_scheduling(name):
    tc = AWAY[name]
    f = fx.Fixtures()
    cfg = fx.test_scheduling_config(disable_away_scheduling=tc.get("no_away", False), disable_gang_away_scheduling=tc.get("no_gang_away", False))
    cfg.well_known_node_types = {"gpu": (tc["wk"],), "large": (Taint("large", "true", "NoSchedule"),)}  # :1041-1044
    node = f.cpu32()  # :1057
    node.taints = node.taints + (tc["node"],)  # :1058-1061
    jobs = [f.job("A", fx.PriorityClass4PreemptibleAway, {"cpu": "1", "memory": "4Gi"})]  # :1064-1068
    if tc.get("gang"):
        # WithGangAnnotationsJobs([job, job])[0] (:1069-1071): the jctx is IN a gang of cardinality 2, and
        # ScheduleManyWithTxn gets a gang context of that one member (:1079-1081)
        jobs = fx.with_gang(jobs + [f.job("A", fx.PriorityClass4PreemptibleAway, {"cpu": "1", "memory": "4Gi"})])
    b, ok, nodes, sched_at, method = _case(cfg, node, jobs)
    assert ok == tc["ok"]  # :1084 / :1104
    if ok:
        assert (nodes == 0).all()  # :1088
        assert (method == abi.METHOD_AWAY).all() and (sched_at == 29000).all()  # :1089-1090
    _round_agrees(b, ok, method, sched_at)
