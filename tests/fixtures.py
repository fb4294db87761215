"""Python restatement of the reference's test fixtures
(internal/scheduler/testfixtures/testfixtures.go:78-135,221-240,615-930,932-1073,1249-1255) so the
reference's table tests can be transcribed one-to-one."""
from __future__ import annotations

import copy
import itertools
import math
from typing import Dict, List, Optional, Sequence

from armada_b200.model import (AwayNodeType, JobSpec, MatchExpression, NodeSpec, PriorityClass, QueueSpec,
                               ResourceType, SchedulingConfig, Taint, Toleration)

PriorityClass0 = "priority-0"
PriorityClass1 = "priority-1"
PriorityClass2 = "priority-2"
PriorityClass2NonPreemptible = "priority-2-non-preemptible"
PriorityClass3 = "priority-3"
PriorityClass4PreemptibleAway = "armada-preemptible-away"  # the reference's names (testfixtures.go:50-58)
PriorityClass5PreemptibleAwayLowPriority = "armada-preemptible-away-lower"
PriorityClass6Preemptible = "armada-preemptible"
PriorityClass7PreemptibleAwayConditional = "armada-preemptible-away-conditional"

TestHostnameLabel = "kubernetes.io/hostname"
ClusterNameLabel = "cluster"
PoolNameLabel = "pool"
NodeTypeLabel = "armadaproject.io/node-type"


def test_priority_classes() -> Dict[str, PriorityClass]:  # testfixtures.go:78-105
    return {
        PriorityClass0: PriorityClass(0, True),
        PriorityClass1: PriorityClass(1, True),
        PriorityClass2: PriorityClass(2, True),
        PriorityClass2NonPreemptible: PriorityClass(2, False),
        PriorityClass3: PriorityClass(3, False),
        PriorityClass4PreemptibleAway: PriorityClass(30000, True, (AwayNodeType(29000, "gpu"), AwayNodeType(29000, "large"))),
        PriorityClass5PreemptibleAwayLowPriority: PriorityClass(30000, True, (AwayNodeType(28000, "gpu"), AwayNodeType(28000, "large"))),
        PriorityClass6Preemptible: PriorityClass(30000, True),
        PriorityClass7PreemptibleAwayConditional: PriorityClass(
            30000, True, (AwayNodeType(29000, "large", (("gpu", (("nvidia.com/gpu", "==", "0"),)),)),)),
    }


TestPriorities = [0, 1, 2, 3, 28000, 29000, 30000]  # testfixtures.go:107


def supported_resource_types():  # GetTestSupportedResourceTypes :1249-1255
    return [ResourceType("memory", "1"), ResourceType("cpu", "1m"), ResourceType("nvidia.com/gpu", "1m")]


def test_resources():  # TestResources :108-112
    return [ResourceType("cpu", "1"), ResourceType("memory", "128Mi"), ResourceType("nvidia.com/gpu", "1")]


def test_scheduling_config(**overrides) -> SchedulingConfig:  # TestSchedulingConfig :221-240
    cfg = SchedulingConfig(
        supported_resource_types=supported_resource_types(),
        indexed_resources=test_resources(),
        priority_classes=test_priority_classes(),
        indexed_taints=["largeJobsOnly", "gpu"],
        indexed_node_labels=["largeJobsOnly", "gpu", ClusterNameLabel, PoolNameLabel, NodeTypeLabel],
        well_known_node_types={
            "gpu": (Taint("gpu", "true", "NoSchedule"),),
            "large": (Taint("largeJobsOnly", "true", "NoSchedule"),),
        },
        drf_resources=["cpu", "memory", "nvidia.com/gpu"],
        enable_prefer_large_job_ordering=True,
    )
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


class Fixtures:
    """Deterministic id / timestamp generators (MockIDProvider / jobTimestamp)."""

    def __init__(self):
        self._node = itertools.count()
        self._job = itertools.count(1)

    # ---- nodes -------------------------------------------------------------------------
    def node(self, resources: Dict[str, str], taints=(), labels=None, node_id: Optional[str] = None) -> NodeSpec:
        i = next(self._node)
        nid = node_id if node_id is not None else f"node-{i:06d}"
        lab = {TestHostnameLabel: nid}
        if labels:
            lab.update(labels)
        return NodeSpec(id=nid, index=i, total=dict(resources), taints=tuple(taints), labels=lab)

    def cpu16(self):
        return self.node({"cpu": "16", "memory": "128Gi"})

    def cpu32(self):  # Test32CpuNode :1030
        return self.node({"cpu": "32", "memory": "256Gi"})

    def tainted_cpu32(self):  # TestTainted32CpuNode :1040
        return self.node({"cpu": "32", "memory": "256Gi"}, taints=(Taint("largeJobsOnly", "true", "NoSchedule"),),
                         labels={"largeJobsOnly": "true"})

    def gpu8(self):  # Test8GpuNode :1060
        return self.node({"cpu": "64", "memory": "1024Gi", "nvidia.com/gpu": "8"}, labels={"gpu": "true"})

    def gpu8_tainted(self):  # WithGpuTaint
        n = self.gpu8()
        n.taints = n.taints + (Taint("gpu", "true", "NoSchedule"),)
        return n

    def n_cpu32(self, n):
        return [self.cpu32() for _ in range(n)]

    def n_tainted_cpu32(self, n):
        return [self.tainted_cpu32() for _ in range(n)]

    def n_gpu8(self, n):
        return [self.gpu8() for _ in range(n)]

    # ---- jobs --------------------------------------------------------------------------
    def job(self, queue: str, pc: str, requests: Dict[str, str], tolerations=(), **kw) -> JobSpec:
        i = next(self._job)
        return JobSpec(id=f"job-{i:09d}", queue=queue, priority_class=pc, requests=dict(requests), queue_priority=1000,
                       submit_time=i, tolerations=tuple(tolerations), **kw)

    def n_jobs(self, queue, pc, n, requests, tolerations=()):
        return [self.job(queue, pc, requests, tolerations) for _ in range(n)]

    def n_1cpu_4gi(self, queue, pc, n):
        return self.n_jobs(queue, pc, n, {"cpu": "1", "memory": "4Gi"})

    def n_1cpu_16gi(self, queue, pc, n):
        return self.n_jobs(queue, pc, n, {"cpu": "1", "memory": "16Gi"})

    def n_16cpu_128gi(self, queue, pc, n):
        return self.n_jobs(queue, pc, n, {"cpu": "16", "memory": "128Gi"})

    def n_16cpu_128gi_large_toleration(self, queue, pc, n):
        return self.n_jobs(queue, pc, n, {"cpu": "16", "memory": "128Gi"}, (Toleration("largeJobsOnly", "", "true"),))

    def n_32cpu_256gi(self, queue, pc, n):
        return self.n_jobs(queue, pc, n, {"cpu": "32", "memory": "256Gi"})

    def n_32cpu_256gi_large_toleration(self, queue, pc, n):
        return self.n_jobs(queue, pc, n, {"cpu": "32", "memory": "256Gi"}, (Toleration("largeJobsOnly", "", "true"),))

    def n_64cpu_512gi(self, queue, pc, n):
        return self.n_jobs(queue, pc, n, {"cpu": "64", "memory": "512Gi"})

    def n_1gpu(self, queue, pc, n):
        return self.n_jobs(queue, pc, n, {"cpu": "8", "memory": "128Gi", "nvidia.com/gpu": "1"}, (Toleration("gpu", "", "true"),))


def with_gang(jobs: Sequence[JobSpec], gang_id: Optional[str] = None) -> List[JobSpec]:
    """WithGangAnnotationsJobs: all jobs form one gang of cardinality len(jobs)."""
    gid = gang_id or f"gang-{jobs[0].id}"
    for j in jobs:
        j.gang_id = gid
        j.gang_cardinality = len(jobs)
    return list(jobs)


def with_node_selector(jobs: Sequence[JobSpec], selector: Dict[str, str]) -> List[JobSpec]:
    for j in jobs:
        j.node_selector = dict(selector)
    return list(jobs)


def with_priority(jobs: Sequence[JobSpec], p: int) -> List[JobSpec]:
    for j in jobs:
        j.queue_priority = p
    return list(jobs)
