"""armada_b200.simulator: the reference's simulator scenarios
(internal/scheduler/simulator/simulator_test.go:34-520) replayed through the event loop, YAML specs of
the reference's schema, parquet sinks.  CPU tests drive the loop with the oracle as the round engine
(host logic only); the GPU tests run the product engine and compare it with the oracle-driven run."""
import os

import pytest

import fixtures as fx
import oracle_lib
from armada_b200 import simulator as sim

HERE = os.path.dirname(os.path.abspath(__file__))
SIMDIR = os.path.join(HERE, "golden", "sim")
MIN = 60 * sim.NS


def oracle_engine(inp):
    return oracle_lib.round_schedule(inp)


def node32(n):  # NodeTemplate32Cpu, test_utils.go:94-104
    return sim.NodeTemplate(n, {"cpu": "32", "memory": "256Gi"})


def jt32(n, job_set, pc, **kw):  # JobTemplate32Cpu, test_utils.go:141-157
    return dict(number=n, job_set=job_set, priority_class_name=pc, requests={"cpu": "32", "memory": "256Gi"},
                runtime=sim.ShiftedExponential(minimum=MIN), **kw)


def workload(*queues):
    qs = []
    for name, templates in queues:
        ts = [sim.JobTemplate(id=t.pop("id", f"{name}-{i}"), queue=name, **t) for i, t in enumerate(templates)]
        qs.append(sim.Queue(name, 1.0, ts))
    return sim.WorkloadSpec("w", qs, random_seed=1)


def cluster(*clusters):
    return sim.ClusterSpec("basic", [sim.Cluster(name, "TestPool", [node32(n)]) for name, n in clusters])


def summary(s):
    """EventSequencesSummary (test_utils.go:228-262) collapsed to (kind, queue) runs."""
    out = []
    for t in s.transitions:
        if out and out[-1][0] == t.kind and out[-1][1] == t.queue:
            out[-1][2] += 1
        else:
            out.append([t.kind, t.queue, 1])
    return [tuple(x) for x in out]


DEFAULT = fx.PriorityClass3  # testfixtures.TestDefaultPriorityClass
SCENARIOS = {
    # simulator_test.go:34-60
    "Two jobs in parallel": (cluster(("TestCluster", 2)), lambda: workload(("A", [jt32(2, "foo", DEFAULT)])), 5,
                             [("submit", "A", 2), ("leased", "A", 2), ("succeeded", "A", 2)]),
    # :61-88
    "Two jobs in sequence": (cluster(("TestCluster", 1)), lambda: workload(("A", [jt32(2, "foo", DEFAULT)])), 5,
                             [("submit", "A", 2), ("leased", "A", 1), ("succeeded", "A", 1), ("leased", "A", 1), ("succeeded", "A", 1)]),
    # :89-112
    "10 jobs in sequence": (cluster(("TestCluster", 1)), lambda: workload(("A", [jt32(10, "foo", DEFAULT)])), 20,
                            [("submit", "A", 10)] + [("leased", "A", 1), ("succeeded", "A", 1)] * 10),
    # :113-150 (two clusters, one pool)
    "Multiple Clusters": (cluster(("ClusterA", 1), ("ClusterB", 1)), lambda: workload(("A", [jt32(2, "foo", DEFAULT)])), 5,
                          [("submit", "A", 2), ("leased", "A", 2), ("succeeded", "A", 2)]),
    # :152-190
    "JobTemplate dependencies": (cluster(("TestCluster", 3)),
                                 lambda: workload(("A", [jt32(2, "foo", DEFAULT, id="jobTemplate"), jt32(1, "foo", DEFAULT, dependencies=["jobTemplate"])])), 5,
                                 [("submit", "A", 2), ("leased", "A", 2), ("succeeded", "A", 2), ("submit", "A", 1), ("leased", "A", 1), ("succeeded", "A", 1)]),
    # :191-232
    "Preemption": (cluster(("TestCluster", 2)),
                   lambda: workload(("A", [jt32(2, "foo", fx.PriorityClass0)]), ("B", [jt32(1, "bar", fx.PriorityClass0, earliest_submit_time=30 * sim.NS)])), 5,
                   [("submit", "A", 2), ("leased", "A", 2), ("submit", "B", 1), ("preempted", "A", 1), ("leased", "B", 1), ("submit", "A", 1),
                    ("succeeded", "A", 1), ("leased", "A", 1), ("succeeded", "B", 1), ("succeeded", "A", 1)]),
    # :233-290: A's job preempts one of C's; the preempted job comes back without pushing anything else out
    "No preemption cascade": (cluster(("TestCluster1", 1), ("TestCluster2", 1), ("TestCluster3", 1)),
                              lambda: workload(("B", [jt32(1, "foo", fx.PriorityClass0)]), ("C", [jt32(2, "foo", fx.PriorityClass0)]),
                                               ("A", [jt32(1, "foo", fx.PriorityClass0, earliest_submit_time=30 * sim.NS)])), 5,
                              [("submit", "B", 1), ("submit", "C", 2), ("leased", "B", 1), ("leased", "C", 2), ("submit", "A", 1), ("preempted", "C", 1),
                               ("leased", "A", 1), ("submit", "C", 1), ("succeeded", "B", 1), ("succeeded", "C", 1), ("leased", "C", 1), ("succeeded", "A", 1),
                               ("succeeded", "C", 1)]),
    # :441-479: Cluster2 is too small for a gang; the uniformity label (the cluster name) keeps every gang on Cluster1
    "Gang Job": (cluster(("Cluster1", 8), ("Cluster2", 1)), lambda: workload(("A", [jt32(16, "foo", DEFAULT, gang_cardinality=8)])), 5,
                 [("submit", "A", 16), ("leased", "A", 8), ("succeeded", "A", 8), ("leased", "A", 8), ("succeeded", "A", 8)]),
    # :480-520
    "Preempted Gang Job": (cluster(("Cluster1", 8)),
                           lambda: workload(("A", [jt32(8, "foo", fx.PriorityClass2, gang_cardinality=8)]),
                                            ("B", [jt32(1, "bar", fx.PriorityClass3, earliest_submit_time=30 * sim.NS)])), 5,
                           [("submit", "A", 8), ("leased", "A", 8), ("submit", "B", 1), ("preempted", "A", 8), ("leased", "B", 1), ("submit", "A", 8),
                            ("succeeded", "B", 1), ("leased", "A", 8), ("succeeded", "A", 8)]),
}


def run_scenario(name, engine):
    cl, wl, minutes, want = SCENARIOS[name]
    s = sim.Simulator(cl, wl(), fx.test_scheduling_config(), engine=engine, hard_termination_minutes=minutes).run()
    return s, want


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_reference_scenarios(name):
    s, want = run_scenario(name, oracle_engine)
    assert summary(s) == want


def test_yaml_specs_and_c1(tmp_path):
    """BASELINE config C1 from YAML files of the reference's schema: 100 nodes x 32 cpu, one queue, 1000 jobs
    of (1 cpu, 10Gi), a 2.5 % per-round limit, 5 minute jobs."""
    out = str(tmp_path / "out")
    s = sim.simulate_files(os.path.join(SIMDIR, "cluster_100x32.yaml"), os.path.join(SIMDIR, "workload_1k.yaml"),
                           os.path.join(SIMDIR, "config_basic.yaml"), output_dir=out, engine=oracle_engine)
    rows = s.sink.job_rows
    assert len(rows) == 1000 and all(r.state == "SUCCEEDED" for r in rows)
    assert all(r.finished_time - r.scheduled_time == 300 for r in rows)
    assert all(r.cpu == 1.0 and r.memory == 10 * 2**30 and r.priority_class == "armada-default" for r in rows)
    # maximumResourceFractionToSchedule 0.025 of 3200 cpu = 80 cpu; CheckRoundConstraints trips once the scheduled resources
    # EXCEED the limit (constraints.go:113-130), i.e. after the 81st job of a cycle; 10 s cycles
    per_cycle = {}
    for r in rows:
        per_cycle[r.scheduled_time] = per_cycle.get(r.scheduled_time, 0) + 1
    assert sorted(per_cycle) == [10 * i for i in range(13)]
    assert [per_cycle[10 * i] for i in range(13)] == [81] * 12 + [28]
    import pyarrow.parquet as pq
    jobs = pq.read_table(os.path.join(out, "jobs.parquet"))
    assert jobs.num_rows == 1000
    assert jobs.column_names == ["queue", "job_set", "job_id", "run_id", "priority_class", "cpu", "memory", "gpu", "ephemeral_storage",
                                 "exit_code", "state", "submitted_time", "scheduled_time", "finished_time"]
    qs = pq.read_table(os.path.join(out, "queue_stats.parquet"))
    assert qs.column_names[:4] == ["ts", "queue", "pool", "fair_share"]
    first = qs.slice(0, 1).to_pylist()[0]
    assert first["queue"] == "A" and first["pool"] == "default" and first["num_scheduled"] == 81 and first["allocated_cpu"] == 81
    assert first["fair_share"] == 1.0


def test_duration_and_spec_parsing():
    assert sim.parse_duration("5m") == 300 * sim.NS and sim.parse_duration("1h30m") == 5400 * sim.NS
    assert sim.parse_duration("500ms") == sim.NS // 2 and sim.parse_duration(None) == 0
    w = sim.workload_spec_from_dict({"queues": [{"name": "A", "weight": 2, "jobTemplates": [
        {"number": 4, "gangCardinality": 2, "repeat": {"numTimes": 3, "period": "1m"}, "requirements": {"resourceRequirements": {"requests": {"cpu": 1}}}}]}]})
    assert w.queues[0].job_templates[0].id == "A-0"
    e = sim.expand_repeating_templates(w)
    assert [t.id for t in e.queues[0].job_templates] == ["A-0-repeat-0", "A-0-repeat-1", "A-0-repeat-2"]
    assert [t.earliest_submit_time for t in e.queues[0].job_templates] == [0, MIN, 2 * MIN]
    with pytest.raises(ValueError):
        sim.workload_spec_from_dict({"queues": [{"name": "A", "weight": 0, "jobTemplates": []}]})
    with pytest.raises(ValueError):
        sim.workload_spec_from_dict({"queues": [{"name": "A", "weight": 1, "jobTemplates": [{"number": 3, "gangCardinality": 2}]}]})


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Preemption", "Preempted Gang Job", "Gang Job", "10 jobs in sequence"])
def test_reference_scenarios_on_the_device(name):
    s, want = run_scenario(name, sim.device_engine(0))
    assert summary(s) == want


@pytest.mark.gpu
def test_c1_on_the_device_matches_the_oracle_driven_run():
    args = (os.path.join(SIMDIR, "cluster_100x32.yaml"), os.path.join(SIMDIR, "workload_1k.yaml"), os.path.join(SIMDIR, "config_basic.yaml"))
    dev = sim.simulate_files(*args, engine=sim.device_engine(0))
    ref = sim.simulate_files(*args, engine=oracle_engine)
    assert dev.rounds == ref.rounds and dev.rounds > 13
    assert dev.sink.job_rows == ref.sink.job_rows
    assert repr(dev.sink.queue_rows) == repr(ref.sink.queue_rows)  # (gpu_share is NaN = 0/0 like the reference's: compare the text)
    assert dev.transitions == ref.transitions


def test_command_line(tmp_path, monkeypatch, capsys):
    """cmd/simulator's flags (cmd/simulator/cmd/root.go:24-33); the product engine is the device — here swapped for the
    oracle so that the host logic runs on a CPU box."""
    monkeypatch.setattr(sim, "device_engine", lambda device=0: oracle_engine)
    out = str(tmp_path / "o")
    args = ["--clusters", os.path.join(SIMDIR, "cluster_100x32.yaml"), "--workloads", os.path.join(SIMDIR, "workload_1k.yaml"),
            "--config", os.path.join(SIMDIR, "config_basic.yaml"), "--outputDir", out, "--hardTerminationMinutes", "2"]
    assert sim.main(args) == 0
    assert "job rows" in capsys.readouterr().out
    assert sorted(os.listdir(out)) == ["jobs.parquet", "queue_stats.parquet"]
    assert sim.main(args) == 1  # exists and is not empty
    assert sim.main(args + ["--overwriteOutputDir"]) == 0


def test_the_product_engine_needs_the_device():
    import ctypes
    from armada_b200 import abi
    try:
        ctypes.CDLL("libcuda.so.1")
        have_driver = True
    except OSError:
        have_driver = False
    if have_driver:
        pytest.skip("a CUDA driver is present")
    with pytest.raises((abi.ArmadaError, OSError)):
        sim.device_engine(0)


def _home_away_scenario():
    """"Home-away preemption" (simulator_test.go:332-440): 32 one-GPU away jobs fill both clusters (16 of them away on the
    tainted whale nodes); two whole-node home jobs arrive, push the 16 away jobs out, run, and the 16 come back."""
    from armada_b200.model import AwayNodeType, PriorityClass, Taint, Toleration
    whale = Taint("gpu-whale", "true", "NoSchedule")
    gpu_node = {"cpu": "128", "memory": "4096Gi", "nvidia.com/gpu": "8"}  # NodeTemplateGpu, test_utils.go:106-116
    cl = sim.ClusterSpec("cluster", [sim.Cluster("WhaleCluster", "TestPool", [sim.NodeTemplate(2, gpu_node, {}, (whale,))]),
                                     sim.Cluster("TestCluster", "TestPool", [sim.NodeTemplate(2, gpu_node)])])
    wl = sim.WorkloadSpec("w", [
        sim.Queue("queue-0", 1.0, [sim.JobTemplate(id="queue-0-template-0", queue="queue-0", number=2, job_set="job-set-0", priority_class_name="armada-preemptible",
                                                   requests=dict(gpu_node), tolerations=(Toleration("gpu-whale", "", "true", "NoSchedule"),),
                                                   earliest_submit_time=MIN, runtime=sim.ShiftedExponential(minimum=5 * MIN))]),
        sim.Queue("queue-1", 1.0, [sim.JobTemplate(id="queue-1-template-0", queue="queue-1", number=32, job_set="job-set-1", priority_class_name="armada-preemptible-away",
                                                   requests={"cpu": "16", "memory": "512Gi", "nvidia.com/gpu": "1"}, runtime=sim.ShiftedExponential(minimum=60 * MIN))])], random_seed=1)
    cfg = fx.test_scheduling_config(priority_classes={"armada-preemptible-away": PriorityClass(30000, True, (AwayNodeType(29000, "gpu-whale"),)),
                                                      "armada-preemptible": PriorityClass(30000, True)},
                                    well_known_node_types={"gpu-whale": (whale,)})
    want = [("submit", "queue-1", 32), ("leased", "queue-1", 32), ("submit", "queue-0", 2), ("preempted", "queue-1", 16), ("leased", "queue-0", 2),
            ("submit", "queue-1", 16), ("succeeded", "queue-0", 2), ("leased", "queue-1", 16), ("succeeded", "queue-1", 32)]
    return cl, wl, cfg, want


def test_home_away_preemption_scenario():
    cl, wl, cfg, want = _home_away_scenario()
    s = sim.Simulator(cl, wl, cfg, engine=oracle_engine, hard_termination_minutes=24 * 60).run()
    assert summary(s) == want


@pytest.mark.gpu
def test_home_away_preemption_scenario_on_the_device():
    cl, wl, cfg, want = _home_away_scenario()
    s = sim.Simulator(cl, wl, cfg, engine=sim.device_engine(0), hard_termination_minutes=24 * 60).run()
    assert summary(s) == want
