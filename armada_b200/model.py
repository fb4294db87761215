"""Host-side mirror of the reference's scheduler-internal types for ONE round, and the
flattening of those types into the SoA `ArmadaRoundInput` the C ABI consumes.

Mirrors (names follow the reference):
  * ResourceListFactory / quantity scaling   internaltypes/resource_list_factory.go:21-109,
                                              internaltypes/quantity_util.go:7-18
  * Node, NodeType                            internaltypes/node.go:26-62, node_type.go:68-124
  * taint / toleration / selector / affinity  nodedb/nodematching.go:127-240,
    matching (static predicates)              kubernetesobjects/taint/taint.go:44-58,
                                              k8s component-helpers v0.32 (restated below)
  * PriorityClass / AwayNodeType              common/types/scheduling.go:56-109
  * per-round / per-queue limits              scheduling/constraints/constraints.go:213-256
The string logic runs on the host; the device only sees dense ids and bitmaps.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi

UNSCHEDULABLE_TAINT_KEY = "armadaproject.io/unschedulable"  # internaltypes/unschedulable.go:7-17
NODE_ID_LABEL = "armadaproject.io/nodeId"
WILDCARD = "*"  # configuration.WildCardWellKnownNodeTypeValue

_SUFFIX = {
    "": Fraction(1), "n": Fraction(1, 10**9), "u": Fraction(1, 10**6), "m": Fraction(1, 1000), "k": Fraction(10**3), "M": Fraction(10**6), "G": Fraction(10**9),
    "T": Fraction(10**12), "P": Fraction(10**15), "Ki": Fraction(2**10), "Mi": Fraction(2**20),
    "Gi": Fraction(2**30), "Ti": Fraction(2**40), "Pi": Fraction(2**50), "Ei": Fraction(2**60), "E": Fraction(10**18),
}


def parse_quantity(q) -> Fraction:
    """k8s resource.Quantity → exact rational (subset: decimal number + SI/binary suffix)."""
    if isinstance(q, (int, Fraction)):
        return Fraction(q)
    s = str(q).strip()
    for suf in ("Ki", "Mi", "Gi", "Ti", "Pi", "Ei", "n", "u", "m", "k", "M", "G", "T", "P", "E"):
        if s.endswith(suf):
            return Fraction(s[: -len(suf)]) * _SUFFIX[suf]
    return Fraction(s)


@dataclass(frozen=True)
class ResourceType:
    name: str
    resolution: str  # e.g. "1m" → scale -3 ; "1" → scale 0


class ResourceListFactory:
    """internaltypes.ResourceListFactory: fixed resource order + per-resource decimal scale."""

    def __init__(self, supported: Sequence[ResourceType]):
        self.names = [r.name for r in supported]
        self.index = {n: i for i, n in enumerate(self.names)}
        # scale = floor(log10(resolution)), resource_list_factory.go:66-71
        self.scales = [int(math.floor(math.log10(float(parse_quantity(r.resolution))))) for r in supported]
        self.D = len(self.names)

    def _scaled(self, name: str, q, round_up: bool) -> int:
        i = self.index[name]
        v = parse_quantity(q) / (Fraction(10) ** self.scales[i])
        return int(math.ceil(v)) if round_up else int(math.floor(v))

    def from_node(self, res: Dict[str, object]) -> np.ndarray:  # FromNodeProto: round DOWN, ignore unknown
        out = np.zeros(self.D, dtype=np.int64)
        for k, v in res.items():
            if k in self.index:
                out[self.index[k]] = self._scaled(k, v, False)
        return out

    def from_job(self, res: Dict[str, object]) -> np.ndarray:  # FromJobResourceListIgnoreUnknown: round UP
        out = np.zeros(self.D, dtype=np.int64)
        for k, v in res.items():
            if k in self.index:
                out[self.index[k]] = self._scaled(k, v, True)
        return out

    def scaled_value(self, name: str, q) -> int:  # Quantity.ScaledValue (rounds up)
        return self._scaled(name, q, True)


@dataclass(frozen=True)
class Taint:
    key: str
    value: str = ""
    effect: str = "NoSchedule"


@dataclass(frozen=True)
class Toleration:
    key: str = ""
    operator: str = ""  # "", "Equal", "Exists"
    value: str = ""
    effect: str = ""


def toleration_tolerates_taint(t: Toleration, taint: Taint) -> bool:
    """v1.Toleration.ToleratesTaint (k8s api core/v1 toleration.go)."""
    if t.effect and t.effect != taint.effect:
        return False
    if t.key and t.key != taint.key:
        return False
    if t.operator in ("", "Equal"):
        return t.value == taint.value
    return t.operator == "Exists"


def find_untolerated(taints: Sequence[Taint], tolerations: Sequence[Toleration]) -> Optional[Taint]:
    """koTaint.FindMatchingUntoleratedTaint: ALL taints must be tolerated, whatever their effect."""
    for taint in taints:
        if not any(toleration_tolerates_taint(t, taint) for t in tolerations):
            return taint
    return None


@dataclass(frozen=True)
class MatchExpression:
    key: str
    operator: str  # In, NotIn, Exists, DoesNotExist, Gt, Lt
    values: Tuple[str, ...] = ()


def _match_expr(labels: Dict[str, str], e: MatchExpression) -> bool:
    has = e.key in labels
    v = labels.get(e.key)
    if e.operator == "In":
        return has and v in e.values
    if e.operator == "NotIn":
        return (not has) or v not in e.values
    if e.operator == "Exists":
        return has
    if e.operator == "DoesNotExist":
        return not has
    if e.operator in ("Gt", "Lt"):
        if not has or len(e.values) != 1:
            return False
        try:
            a, b = int(v), int(e.values[0])
        except ValueError:
            return False
        return a > b if e.operator == "Gt" else a < b
    return False


def match_node_selector_terms(labels: Dict[str, str], terms: Sequence[Tuple[MatchExpression, ...]]) -> bool:
    """corev1.MatchNodeSelectorTerms: terms are ORed, expressions ANDed, empty term matches nothing."""
    for term in terms:
        if len(term) == 0:
            continue
        if all(_match_expr(labels, e) for e in term):
            return True
    return False


@dataclass
class AwayNodeType:
    priority: int
    well_known_node_type: str = ""
    # extra (name, [(resource, op, value)]) entries: types.AwayTypeEntry
    node_types: Tuple[Tuple[str, Tuple[Tuple[str, str, str], ...]], ...] = ()


@dataclass
class PriorityClass:
    priority: int
    preemptible: bool
    away_node_types: Tuple[AwayNodeType, ...] = ()
    # MaximumResourceFractionPerQueue (types/scheduling.go:62-68)
    maximum_resource_fraction_per_queue: Dict[str, float] = field(default_factory=dict)


@dataclass(frozen=True)
class FloatingResource:
    name: str
    resolution: str = "1"
    quantity: Optional[object] = None  # the pool's total; None = this pool is not listed for the resource


@dataclass
class SchedulingConfig:
    """The hot-path-relevant subset of configuration.SchedulingConfig (configuration.go:186-361)."""
    supported_resource_types: Sequence[ResourceType]
    indexed_resources: Sequence[ResourceType]  # resolution in quantity units, e.g. cpu "1", memory "128Mi"
    priority_classes: Dict[str, PriorityClass]
    indexed_taints: Optional[Sequence[str]] = None  # None ⇒ all taints indexed
    indexed_node_labels: Sequence[str] = ()
    well_known_node_types: Dict[str, Tuple[Taint, ...]] = field(default_factory=dict)
    drf_resources: Sequence[str] = ()  # DominantResourceFairnessResourcesToConsider
    drf_multipliers: Optional[Dict[str, float]] = None  # Experimental…ResourcesToConsider
    protected_fraction_of_fair_share: float = 0.0
    protect_uncapped_adjusted_fair_share: bool = False
    max_queue_lookback: int = 0
    maximum_resource_fraction_to_schedule: Optional[Dict[str, float]] = None
    maximum_scheduling_rate: float = math.inf
    maximum_scheduling_burst: int = 2**62
    maximum_per_queue_scheduling_rate: float = math.inf
    maximum_per_queue_scheduling_burst: int = 2**62
    enable_prefer_large_job_ordering: bool = True
    disable_home_scheduling: bool = False
    disable_away_scheduling: bool = False
    disable_gang_away_scheduling: bool = False
    disallowed_resources: Sequence[str] = ()
    # FloatingResources of THIS pool (configuration.FloatingResourceConfig): resources no node holds, limited per pool
    floating_resources: Sequence["FloatingResource"] = ()

    def factory(self) -> ResourceListFactory:
        # NewResourceListFactory: the supported (Kubernetes) types, then the floating ones (resource_list_factory.go:41-53)
        return ResourceListFactory(list(self.supported_resource_types) + [ResourceType(f.name, f.resolution) for f in self.floating_resources])

    def allowed_priorities(self) -> List[int]:
        """types.AllowedPriorities (common/types/scheduling.go:99-109): PC + away priorities, sorted, unique."""
        ps = set()
        for pc in self.priority_classes.values():
            ps.add(pc.priority)
            for a in pc.away_node_types:
                ps.add(a.priority)
        return sorted(ps)


@dataclass
class NodeSpec:
    id: str
    index: int
    total: Dict[str, object]
    taints: Tuple[Taint, ...] = ()
    labels: Dict[str, str] = field(default_factory=dict)
    allocatable: Optional[Dict[str, object]] = None  # default: total
    unschedulable: bool = False
    over_allocated: bool = False


@dataclass
class JobSpec:
    id: str
    queue: str
    priority_class: str
    requests: Dict[str, object]
    queue_priority: int = 0
    submit_time: int = 0
    tolerations: Tuple[Toleration, ...] = ()
    node_selector: Dict[str, str] = field(default_factory=dict)
    affinity: Optional[Tuple[Tuple[MatchExpression, ...], ...]] = None  # required node-affinity terms
    gang_id: Optional[str] = None
    gang_cardinality: int = 1
    gang_node_uniformity_label: Optional[str] = None  # GangInfo.NodeUniformity() (jobdb/gang.go)
    node: Optional[str] = None  # id of the node the job runs on (None = queued)
    scheduled_at_priority: Optional[int] = None
    active_run_timestamp: int = 0


@dataclass
class QueueSpec:
    name: str
    priority_factor: float = 1.0
    cordoned: bool = False
    allocated_by_pc: Dict[str, np.ndarray] = field(default_factory=dict)  # pc name -> int64[D]
    demand: Optional[np.ndarray] = None  # int64[D]
    constrained_demand: Optional[np.ndarray] = None
    short_job_penalty: Optional[np.ndarray] = None
    limiter_tokens: Optional[float] = None  # None ⇒ burst (fresh limiter)
    # per-queue overrides of MaximumResourceFraction by priority class name
    resource_limits_by_pc: Dict[str, Dict[str, float]] = field(default_factory=dict)


def multiply_resource(res: int, m: float) -> int:
    """multiplyResource, internaltypes/resource_list.go:312-331."""
    if m == 1.0:
        return int(res)
    if math.isinf(m):
        return (2**63 - 1) if ((m < 0) == (res < 0)) else -(2**63)
    v = float(res) * m
    if v >= 9.223372036854775807e18:  # Go int64(float) overflow is implementation-defined; saturate
        return 2**63 - 1
    if v <= -9.223372036854775808e18:
        return -(2**63)
    return int(v)


class RoundInputBuilder:
    """Flattens (config, nodes, jobs, queues) into an ArmadaRoundInput.  Keeps the numpy arrays alive."""

    def __init__(self, cfg: SchedulingConfig, nodes: Sequence[NodeSpec], jobs: Sequence[JobSpec],
                 queues: Sequence[QueueSpec], total_resources: Optional[np.ndarray] = None,
                 queued_order: Optional[Dict[str, List[str]]] = None, global_limiter_tokens: Optional[float] = None):
        self.cfg = cfg
        self.factory = cfg.factory()
        self.nodes = list(nodes)
        self.jobs = list(jobs)
        self.queues = sorted(queues, key=lambda q: q.name)  # index order == name order
        self.queue_index = {q.name: i for i, q in enumerate(self.queues)}
        self.node_pos = {n.id: i for i, n in enumerate(self.nodes)}
        self.job_pos = {j.id: i for i, j in enumerate(self.jobs)}
        self._keep: List[object] = []
        self.pc_names = sorted(cfg.priority_classes.keys())
        self.pc_index = {n: i for i, n in enumerate(self.pc_names)}
        self._build(total_resources, queued_order, global_limiter_tokens)

    # -- helpers ---------------------------------------------------------------------------
    def _arr(self, a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        self._keep.append(a)
        return a

    def _ptr(self, a, ctype):
        if a is None:
            return None
        return a.ctypes.data_as(C.POINTER(ctype))

    def _node_taints(self, n: NodeSpec) -> Tuple[Taint, ...]:
        t = tuple(n.taints)
        if n.unschedulable:  # CreateNodeAndType, node.go:120-122
            t = t + (Taint(UNSCHEDULABLE_TAINT_KEY, "true", "NoSchedule"),)
        return t

    def _node_labels(self, n: NodeSpec) -> Dict[str, str]:
        labels = dict(n.labels)
        labels[NODE_ID_LABEL] = n.id
        return labels

    def _away_tolerations(self, pc: PriorityClass, k: int, req: np.ndarray) -> Tuple[Toleration, ...]:
        """getEffectiveAwayNodeTaints + toleration synthesis, nodedb.go:532-598."""
        away = pc.away_node_types[k]
        taints: List[Taint] = []
        if away.well_known_node_type:
            taints += list(self.cfg.well_known_node_types.get(away.well_known_node_type, ()))
        for name, conditions in away.node_types:
            ok = True
            for (resource, op, value) in conditions:  # matchesCondition :514-530 (first condition decides)
                jv = int(parse_quantity(0))
                if resource in self.factory.index:
                    i = self.factory.index[resource]
                    # GetByNameZeroIfMissing(...).Value(): quantity rounded up to an integer
                    jv = int(math.ceil(Fraction(int(req[i])) * (Fraction(10) ** self.factory.scales[i])))
                v = int(math.ceil(parse_quantity(value)))
                ok = {">": jv > v, "<": jv < v, "==": jv == v}[op]
                break
            if ok:
                taints += list(self.cfg.well_known_node_types[name])
        out = []
        for t in taints:
            if t.value == WILDCARD:
                out.append(Toleration(key=t.key, operator="Exists", effect=t.effect))
            else:
                out.append(Toleration(key=t.key, value=t.value, effect=t.effect))
        return tuple(out)

    # -- main ------------------------------------------------------------------------------
    def _build(self, total_resources, queued_order, global_limiter_tokens):
        cfg, f = self.cfg, self.factory
        D = f.D
        inp = abi.RoundInput()
        inp.abi_version = abi.ABI_VERSION
        inp.num_resources = D
        R = len(cfg.indexed_resources)
        inp.num_indexed = R
        for i, r in enumerate(cfg.indexed_resources):
            inp.indexed_resource[i] = f.index[r.name]
            inp.indexed_resolution[i] = f.scaled_value(r.name, r.resolution)  # makeIndexedResourceResolution
        prios = [-1] + cfg.allowed_priorities()
        inp.num_priorities = len(prios)
        for i, p in enumerate(prios):
            inp.priorities[i] = p
        self.priorities = prios
        well_known_names = sorted(cfg.well_known_node_types.keys())
        inp.num_priority_classes = len(self.pc_names)
        for i, name in enumerate(self.pc_names):
            pc = cfg.priority_classes[name]
            s = inp.priority_classes[i]
            s.priority = pc.priority
            s.preemptible = 1 if pc.preemptible else 0
            s.num_away = len(pc.away_node_types)
            for k, a in enumerate(pc.away_node_types):
                s.away_priority[k] = a.priority
                # (a name without a WellKnownNodeTypes entry adds no taints here; the reference fails the away attempt that reaches it)
                s.away_well_known[k] = well_known_names.index(a.well_known_node_type) if a.well_known_node_type in well_known_names else abi.NONE

        # ---- nodes ----
        N = len(self.nodes)
        inp.num_nodes = N
        node_total = np.zeros((D, N), dtype=np.int64)
        node_alloc = np.zeros((D, N), dtype=np.int64)
        for i, n in enumerate(self.nodes):
            node_total[:, i] = f.from_node(n.total)
            node_alloc[:, i] = f.from_node(n.allocatable if n.allocatable is not None else n.total)
        ids_sorted = sorted(range(N), key=lambda i: self.nodes[i].id)
        id_rank = np.zeros(N, dtype=np.uint32)
        for r, i in enumerate(ids_sorted):
            id_rank[i] = r
        indexed_taints = None if cfg.indexed_taints is None else set(cfg.indexed_taints)
        indexed_labels = set(cfg.indexed_node_labels)

        # ---- job classes / static rows ----
        rows: Dict[object, int] = {}
        row_specs: List[Tuple[Tuple[Toleration, ...], Tuple[Tuple[str, str], ...], object]] = []

        def row_of(tolerations, selector, affinity) -> int:
            key = (tuple(sorted(tolerations, key=repr)), tuple(sorted(selector.items())), affinity)
            if key not in rows:
                rows[key] = len(row_specs)
                row_specs.append((tuple(tolerations), tuple(sorted(selector.items())), affinity))
            return rows[key]

        classes: Dict[object, int] = {}
        class_req: List[np.ndarray] = []
        class_pc: List[int] = []
        class_row: List[int] = []
        class_away: List[List[int]] = []
        class_meta: List[Tuple[Tuple[Toleration, ...], Dict[str, str], object, List[Tuple[Toleration, ...]]]] = []
        job_class = np.zeros(len(self.jobs), dtype=np.uint32)
        for ji, j in enumerate(self.jobs):
            req = f.from_job(j.requests)
            key = (tuple(sorted(j.tolerations, key=repr)), tuple(sorted(j.node_selector.items())), j.affinity,
                   tuple(int(x) for x in req), j.priority_class)
            if key not in classes:
                classes[key] = len(class_req)
                class_req.append(req)
                pc = cfg.priority_classes[j.priority_class]
                class_pc.append(self.pc_index[j.priority_class])
                class_row.append(row_of(j.tolerations, j.node_selector, j.affinity))
                aw = [abi.NONE] * abi.MAX_AWAY
                extras: List[Tuple[Toleration, ...]] = []
                for k in range(len(pc.away_node_types)):
                    extra = self._away_tolerations(pc, k, req)
                    extras.append(extra)
                    if extra:
                        aw[k] = row_of(tuple(j.tolerations) + extra, j.node_selector, j.affinity)
                class_away.append(aw)
                class_meta.append((tuple(j.tolerations), dict(j.node_selector), j.affinity, extras))
            job_class[ji] = classes[key]
        # ---- gang node uniformity (gang_scheduler.go:154-223): value slots per label, rows per (class, slot) ----
        uni_labels: Dict[str, int] = {}
        uni_values: List[List[str]] = []
        uni_classes: Dict[int, set] = {}
        for ji, j in enumerate(self.jobs):
            lab = j.gang_node_uniformity_label
            if not lab or j.gang_id is None or j.gang_cardinality <= 1 or lab not in indexed_labels:
                continue
            if lab not in uni_labels:  # nodeDb.IndexedNodeLabelValues(label) minus "" (:181-190)
                uni_labels[lab] = len(uni_values)
                uni_values.append(sorted({self._node_labels(n)[lab] for n in self.nodes if self._node_labels(n).get(lab)}))
            uni_classes.setdefault(uni_labels[lab], set()).add(int(job_class[ji]))
        uni_start = [0]
        for vals in uni_values:
            uni_start.append(uni_start[-1] + len(vals))
        Vn = uni_start[-1]
        class_uni = np.full((max(1, len(class_req)), max(1, Vn), 1 + abi.MAX_AWAY), abi.NONE, dtype=np.uint32)
        for lab, li in uni_labels.items():
            for c in uni_classes.get(li, ()):
                tol, sel, aff, extras = class_meta[c]
                for vi, val in enumerate(uni_values[li]):
                    sel2 = dict(sel)
                    sel2[lab] = val  # jctx.AddNodeSelector
                    class_uni[c, uni_start[li] + vi, 0] = row_of(tol, sel2, aff)
                    for k, extra in enumerate(extras):
                        if extra:
                            class_uni[c, uni_start[li] + vi, 1 + k] = row_of(tol + extra, sel2, aff)
        self.uni_labels, self.uni_values, self.uni_start = uni_labels, uni_values, uni_start
        Cn = max(1, len(class_req))
        if not class_req:  # keep arrays non-empty for the C side
            class_req, class_pc, class_row, class_away = [np.zeros(D, np.int64)], [0], [row_of((), {}, None)], [[abi.NONE] * abi.MAX_AWAY]
        inp.num_classes = len(class_req)

        # relevant label keys: only labels some row looks at distinguish static classes
        rel_keys = set()
        for (_, selector, affinity) in row_specs:
            rel_keys.update(k for k, _ in selector)
            if affinity:
                for term in affinity:
                    rel_keys.update(e.key for e in term)
        static_classes: Dict[object, int] = {}
        static_specs: List[Tuple[Tuple[Taint, ...], Dict[str, str]]] = []
        node_static = np.zeros(N, dtype=np.uint32)
        types: Dict[object, int] = {}
        type_specs: List[Tuple[Tuple[Taint, ...], Dict[str, str], set]] = []
        node_type = np.zeros(N, dtype=np.uint32)
        node_flags = np.zeros(N, dtype=np.uint8)
        for i, n in enumerate(self.nodes):
            taints = self._node_taints(n)
            labels = self._node_labels(n)
            skey = (tuple(sorted(taints, key=repr)), tuple(sorted((k, v) for k, v in labels.items() if k in rel_keys)))
            if skey not in static_classes:
                static_classes[skey] = len(static_specs)
                static_specs.append((taints, labels))
            node_static[i] = static_classes[skey]
            # NewNodeType, node_type.go:68-124
            ttaints = tuple(t for t in taints if indexed_taints is None or t.key in indexed_taints)
            tlabels = {k: v for k, v in labels.items() if k in indexed_labels}
            unset = set(k for k in indexed_labels if k not in tlabels)
            tkey = (tuple(sorted(ttaints, key=repr)), tuple(sorted(tlabels.items())), tuple(sorted(unset)))
            forced = getattr(n, "forced_type", None)  # tests only: WithNodeTypeNodes
            if forced is not None:
                tkey = ("forced", forced)
            if tkey not in types:
                types[tkey] = len(type_specs)
                type_specs.append((ttaints, tlabels, unset))
            node_type[i] = types[tkey]
            node_flags[i] = (abi.NODE_UNSCHEDULABLE if n.unschedulable else 0) | (abi.NODE_OVERALLOCATED if n.over_allocated else 0)
        S, T = max(1, len(static_specs)), max(1, len(type_specs))
        inp.num_static_classes, inp.num_node_types = S, T
        nrows = len(row_specs)
        inp.num_static_rows = nrows
        sw, tw = (S + 31) // 32, (T + 31) // 32
        static_match = np.zeros((nrows, sw), dtype=np.uint32)
        type_match = np.zeros((nrows, tw), dtype=np.uint32)
        for r, (tolerations, selector, affinity) in enumerate(row_specs):
            for s, (taints, labels) in enumerate(static_specs):
                ok = find_untolerated(taints, tolerations) is None  # NodeTolerationRequirementsMet
                ok = ok and all(labels.get(k) == v for k, v in selector)  # NodeSelectorRequirementsMet(node, nil)
                if ok and affinity is not None:
                    ok = match_node_selector_terms(labels, affinity)  # NodeAffinityRequirementsMet
                if ok:
                    static_match[r, s >> 5] |= np.uint32(1 << (s & 31))
            for t, (ttaints, tlabels, unset) in enumerate(type_specs):
                ok = find_untolerated(ttaints, tolerations) is None  # TolerationRequirementsMet(nodeType)
                if ok:
                    for k, v in selector:  # NodeSelectorRequirementsMet(type labels, unsetIndexedLabels)
                        if k in tlabels:
                            if tlabels[k] != v:
                                ok = False
                                break
                        elif k in unset:
                            ok = False
                            break
                if ok:
                    type_match[r, t >> 5] |= np.uint32(1 << (t & 31))
        self.row_specs, self.static_specs, self.type_specs = row_specs, static_specs, type_specs
        self.type_keys = list(types.keys())

        a = self._arr
        self.node_index = a([n.index for n in self.nodes] or [0], np.uint64)
        self.node_id_rank = a(id_rank if N else [0], np.uint32)
        self.node_type = a(node_type if N else [0], np.uint32)
        self.node_static = a(node_static if N else [0], np.uint32)
        self.node_flags = a(node_flags if N else [0], np.uint8)
        self.node_total = a(node_total if N else np.zeros((D, 1)), np.int64)
        self.node_alloc = a(node_alloc if N else np.zeros((D, 1)), np.int64)
        inp.node_index = self._ptr(self.node_index, C.c_uint64)
        inp.node_id_rank = self._ptr(self.node_id_rank, C.c_uint32)
        inp.node_type = self._ptr(self.node_type, C.c_uint32)
        inp.node_static_class = self._ptr(self.node_static, C.c_uint32)
        inp.node_flags = self._ptr(self.node_flags, C.c_uint8)
        inp.node_total = self._ptr(self.node_total, C.c_int64)
        inp.node_allocatable = self._ptr(self.node_alloc, C.c_int64)

        self.class_request = a(np.stack(class_req), np.int64)
        self.class_pc = a(class_pc, np.uint32)
        self.class_row = a(class_row, np.uint32)
        self.class_away = a(class_away, np.uint32)
        self.class_key_valid = a(np.ones(len(class_req)), np.uint8)
        self.static_match = a(static_match, np.uint32)
        self.type_match = a(type_match, np.uint32)
        inp.class_request = self._ptr(self.class_request, C.c_int64)
        inp.class_pc = self._ptr(self.class_pc, C.c_uint32)
        inp.class_static_row = self._ptr(self.class_row, C.c_uint32)
        inp.class_away_row = self._ptr(self.class_away, C.c_uint32)
        inp.class_key_valid = self._ptr(self.class_key_valid, C.c_uint8)
        inp.static_match = self._ptr(self.static_match, C.c_uint32)
        inp.type_match = self._ptr(self.type_match, C.c_uint32)

        # ---- jobs ----
        J = len(self.jobs)
        inp.num_jobs = J
        gangs: Dict[Tuple[str, str], int] = {}
        gang_card: List[int] = []
        job_gang = np.full(max(J, 1), abi.NONE, dtype=np.uint32)
        job_queue = np.full(max(J, 1), abi.NONE, dtype=np.uint32)
        job_node = np.full(max(J, 1), abi.NONE, dtype=np.uint32)
        job_sap = np.full(max(J, 1), abi.NO_PRIORITY, dtype=np.int32)
        job_qp = np.zeros(max(J, 1), dtype=np.uint32)
        job_st = np.zeros(max(J, 1), dtype=np.int64)
        job_art = np.zeros(max(J, 1), dtype=np.int64)
        jid_sorted = sorted(range(J), key=lambda i: self.jobs[i].id)
        job_id_rank = np.zeros(max(J, 1), dtype=np.uint32)
        for r, i in enumerate(jid_sorted):
            job_id_rank[i] = r
        for ji, j in enumerate(self.jobs):
            if j.queue in self.queue_index:
                job_queue[ji] = self.queue_index[j.queue]
            if j.gang_id is not None and j.gang_cardinality > 1:  # GangInfo.IsGang
                gk = (j.queue, j.gang_id)
                if gk not in gangs:
                    gangs[gk] = len(gang_card)
                    gang_card.append(j.gang_cardinality)
                job_gang[ji] = gangs[gk]
            if j.node is not None:
                job_node[ji] = self.node_pos[j.node]
                if j.scheduled_at_priority is not None:
                    job_sap[ji] = j.scheduled_at_priority
            job_qp[ji] = j.queue_priority
            job_st[ji] = j.submit_time
            job_art[ji] = j.active_run_timestamp
        inp.num_gangs = len(gang_card)
        self.job_class = a(job_class if J else [0], np.uint32)
        self.job_queue = a(job_queue, np.uint32)
        self.job_qp = a(job_qp, np.uint32)
        self.job_st = a(job_st, np.int64)
        self.job_id_rank = a(job_id_rank, np.uint32)
        self.job_gang = a(job_gang, np.uint32)
        self.job_node = a(job_node, np.uint32)
        self.job_sap = a(job_sap, np.int32)
        self.job_art = a(job_art, np.int64)
        self.gang_card = a(gang_card or [0], np.uint32)
        inp.job_class = self._ptr(self.job_class, C.c_uint32)
        inp.job_queue = self._ptr(self.job_queue, C.c_uint32)
        inp.job_queue_priority = self._ptr(self.job_qp, C.c_uint32)
        inp.job_submit_time = self._ptr(self.job_st, C.c_int64)
        inp.job_id_rank = self._ptr(self.job_id_rank, C.c_uint32)
        inp.job_gang = self._ptr(self.job_gang, C.c_uint32)
        inp.job_node = self._ptr(self.job_node, C.c_uint32)
        inp.job_scheduled_at_priority = self._ptr(self.job_sap, C.c_int32)
        inp.job_active_run_timestamp = self._ptr(self.job_art, C.c_int64)
        inp.gang_cardinality = self._ptr(self.gang_card, C.c_uint32)
        gang_label = np.full(max(1, len(gang_card)), abi.NONE, dtype=np.uint32)
        for ji, j in enumerate(self.jobs):
            if job_gang[ji] != abi.NONE and j.gang_node_uniformity_label:
                lab = j.gang_node_uniformity_label
                gang_label[job_gang[ji]] = self.uni_labels[lab] if lab in indexed_labels else abi.LABEL_NOT_INDEXED
        self.gang_label = a(gang_label, np.uint32)
        self.uni_start_arr = a(self.uni_start, np.uint32)
        self.class_uni = a(class_uni, np.uint32)
        if (gang_label != abi.NONE).any():
            inp.gang_uniformity_label = self._ptr(self.gang_label, C.c_uint32)
            inp.num_uniformity_labels = len(self.uni_values)
            inp.uniformity_value_start = self._ptr(self.uni_start_arr, C.c_uint32)
            inp.class_uniformity_row = self._ptr(self.class_uni, C.c_uint32)
        # floating resources (floatingresources/floating_resource_types.go:19-37)
        fmask = 0
        for fr in cfg.floating_resources:
            d = f.index[fr.name]
            fmask |= 1 << d
            if fr.quantity is not None:
                inp.floating_limits_configured = 1
                inp.floating_limit[d] = f.scaled_value(fr.name, fr.quantity)
        inp.floating_resource_mask = fmask

        # ---- scheduling context scalars ----
        if total_resources is None:  # nodeDb.TotalKubernetesResources(): Σ allocatable
            total_resources = node_alloc.sum(axis=1) if N else np.zeros(D, np.int64)
        self.total_resources = np.asarray(total_resources, dtype=np.int64)
        for d in range(D):
            inp.total_resources[d] = int(self.total_resources[d])
        mult = {}
        if cfg.drf_multipliers:
            mult = {k: (v if v > 0 else 1.0) for k, v in cfg.drf_multipliers.items()}
        else:
            mult = {k: 1.0 for k in cfg.drf_resources}
        for d, name in enumerate(f.names):
            inp.drf_multipliers[d] = float(mult.get(name, 0.0))
        inp.protected_fraction_of_fair_share = cfg.protected_fraction_of_fair_share
        inp.protect_uncapped_adjusted_fair_share = int(cfg.protect_uncapped_adjusted_fair_share)
        inp.prefer_large_job_ordering = int(cfg.enable_prefer_large_job_ordering)
        inp.disable_home_scheduling = int(cfg.disable_home_scheduling)
        inp.disable_away_scheduling = int(cfg.disable_away_scheduling)
        inp.disable_gang_away_scheduling = int(cfg.disable_gang_away_scheduling)
        inp.max_queue_lookback = cfg.max_queue_lookback
        mask = 0
        for r in cfg.disallowed_resources:
            if r in f.index:
                mask |= 1 << f.index[r]
        inp.disallowed_resource_mask = mask
        # calculatePerRoundLimits, constraints.go:213-229 (missing fractions default to +Inf)
        inp.has_round_limit = 1
        fr = cfg.maximum_resource_fraction_to_schedule or {}
        for d, name in enumerate(f.names):
            inp.max_resources_to_schedule[d] = multiply_resource(int(self.total_resources[d]), fr.get(name, math.inf))
        # rate limiters (golang.org/x/time/rate): fresh limiter ⇒ tokens == burst
        inp.global_limiter_is_inf = int(math.isinf(cfg.maximum_scheduling_rate))
        inp.global_limiter_burst = min(cfg.maximum_scheduling_burst, 2**62)
        inp.global_limiter_tokens = float(inp.global_limiter_burst) if global_limiter_tokens is None else global_limiter_tokens

        # ---- queues ----
        Qn = len(self.queues)
        inp.num_queues = Qn
        PCn = len(self.pc_names)
        qw = np.zeros(max(Qn, 1), dtype=np.float64)
        qc = np.zeros(max(Qn, 1), dtype=np.uint8)
        qa = np.zeros((max(Qn, 1), PCn, D), dtype=np.int64)
        qd = np.zeros((max(Qn, 1), D), dtype=np.int64)
        qcd = np.zeros((max(Qn, 1), D), dtype=np.int64)
        qp = np.zeros((max(Qn, 1), D), dtype=np.int64)
        qhl = np.zeros((max(Qn, 1), PCn), dtype=np.uint8)
        ql = np.zeros((max(Qn, 1), PCn, D), dtype=np.int64)
        qt = np.zeros(max(Qn, 1), dtype=np.float64)
        qb = np.zeros(max(Qn, 1), dtype=np.int64)
        qi = np.zeros(max(Qn, 1), dtype=np.uint8)
        for i, q in enumerate(self.queues):
            qw[i] = 1.0 / q.priority_factor if q.priority_factor > 0 else 1.0
            qc[i] = int(q.cordoned)
            for pcn, rl in q.allocated_by_pc.items():
                qa[i, self.pc_index[pcn]] = rl
            if q.demand is not None:
                qd[i] = q.demand
            if q.constrained_demand is not None:
                qcd[i] = q.constrained_demand
            elif q.demand is not None:
                qcd[i] = q.demand
            if q.short_job_penalty is not None:
                qp[i] = q.short_job_penalty
            # calculatePerQueueLimits, constraints.go:231-256
            for pcn, pc in cfg.priority_classes.items():
                fractions = dict(pc.maximum_resource_fraction_per_queue)
                fractions.update(q.resource_limits_by_pc.get(pcn, {}))
                pi = self.pc_index[pcn]
                qhl[i, pi] = 1
                for d, name in enumerate(f.names):
                    ql[i, pi, d] = multiply_resource(int(self.total_resources[d]), fractions.get(name, math.inf))
            qb[i] = min(cfg.maximum_per_queue_scheduling_burst, 2**62)
            qt[i] = float(qb[i]) if q.limiter_tokens is None else q.limiter_tokens
            qi[i] = int(math.isinf(cfg.maximum_per_queue_scheduling_rate))
        self.qw, self.qc, self.qa, self.qd, self.qcd, self.qp = a(qw, np.float64), a(qc, np.uint8), a(qa, np.int64), a(qd, np.int64), a(qcd, np.int64), a(qp, np.int64)
        self.qhl, self.ql, self.qt, self.qb, self.qi = a(qhl, np.uint8), a(ql, np.int64), a(qt, np.float64), a(qb, np.int64), a(qi, np.uint8)
        inp.queue_weight = self._ptr(self.qw, C.c_double)
        inp.queue_cordoned = self._ptr(self.qc, C.c_uint8)
        inp.queue_allocated_by_pc = self._ptr(self.qa, C.c_int64)
        inp.queue_demand = self._ptr(self.qd, C.c_int64)
        inp.queue_constrained_demand = self._ptr(self.qcd, C.c_int64)
        inp.queue_short_job_penalty = self._ptr(self.qp, C.c_int64)
        inp.queue_has_limit = self._ptr(self.qhl, C.c_uint8)
        inp.queue_limit = self._ptr(self.ql, C.c_int64)
        inp.queue_limiter_tokens = self._ptr(self.qt, C.c_double)
        inp.queue_limiter_burst = self._ptr(self.qb, C.c_int64)
        inp.queue_limiter_is_inf = self._ptr(self.qi, C.c_uint8)

        if queued_order is not None:
            start = [0]
            order: List[int] = []
            for q in self.queues:
                order += [self.job_pos[jid] for jid in queued_order.get(q.name, [])]
                start.append(len(order))
            self.queued_start = a(start, np.uint32)
            self.queued_order = a(order or [0], np.uint32)
            inp.queued_start = self._ptr(self.queued_start, C.c_uint32)
            inp.queued_order = self._ptr(self.queued_order, C.c_uint32)
        inp._keepalive = self._keep
        self.input = inp


class RoundResult:
    """Caller-allocated ArmadaRoundOutput + numpy views."""

    # ask for job_seq_first_pass / job_reason_first_pass (the inputs of `queue_stats`); off by default: 5 bytes per job
    # more to bring back per round.  tests/conftest.py switches it on for every test.
    FIRST_PASS_DEFAULT = False

    def __init__(self, inp: abi.RoundInput, first_pass: Optional[bool] = None):
        J, N, Q = max(inp.num_jobs, 1), max(inp.num_nodes, 1), max(inp.num_queues, 1)
        D, PL, PC = inp.num_resources, inp.num_priorities, inp.num_priority_classes
        self.job_state = np.zeros(J, np.uint8)
        self.job_node = np.full(J, abi.NONE, np.uint32)
        self.job_scheduled_at_priority = np.zeros(J, np.int32)
        self.job_preempted_at_priority = np.zeros(J, np.int32)
        self.job_method = np.zeros(J, np.uint8)
        self.job_reason = np.zeros(J, np.uint8)
        self.job_seq = np.zeros(J, np.uint32)
        self.node_alloc = np.zeros((PL, D, N), np.int64)
        self.queue_allocated = np.zeros((Q, D), np.int64)
        self.queue_allocated_by_pc = np.zeros((Q, PC, D), np.int64)
        self.queue_fair_share = np.zeros((Q, 3), np.float64)
        self.scheduled_resources = np.zeros(D, np.int64)
        self.evicted_resources = np.zeros(D, np.int64)
        self.job_excluded_nodes = np.zeros((J, abi.EXCLUDED_KINDS), np.uint32)
        self.job_seq_first_pass = np.zeros(J, np.uint32)
        self.job_reason_first_pass = np.zeros(J, np.uint8)
        o = abi.RoundOutput()
        o.job_state = self.job_state.ctypes.data_as(abi.u8p)
        o.job_node = self.job_node.ctypes.data_as(abi.u32p)
        o.job_scheduled_at_priority = self.job_scheduled_at_priority.ctypes.data_as(abi.i32p)
        o.job_preempted_at_priority = self.job_preempted_at_priority.ctypes.data_as(abi.i32p)
        o.job_method = self.job_method.ctypes.data_as(abi.u8p)
        o.job_reason = self.job_reason.ctypes.data_as(abi.u8p)
        o.job_seq = self.job_seq.ctypes.data_as(abi.u32p)
        o.node_alloc = self.node_alloc.ctypes.data_as(abi.i64p)
        o.queue_allocated = self.queue_allocated.ctypes.data_as(abi.i64p)
        o.queue_allocated_by_pc = self.queue_allocated_by_pc.ctypes.data_as(abi.i64p)
        o.queue_fair_share = self.queue_fair_share.ctypes.data_as(abi.f64p)
        o.scheduled_resources = self.scheduled_resources.ctypes.data_as(abi.i64p)
        o.evicted_resources = self.evicted_resources.ctypes.data_as(abi.i64p)
        if inp.collect_excluded_nodes:
            o.job_excluded_nodes = self.job_excluded_nodes.ctypes.data_as(abi.u32p)
        self.first_pass = RoundResult.FIRST_PASS_DEFAULT if first_pass is None else first_pass
        if self.first_pass:
            o.job_seq_first_pass = self.job_seq_first_pass.ctypes.data_as(abi.u32p)
            o.job_reason_first_pass = self.job_reason_first_pass.ctypes.data_as(abi.u8p)
        self.out = o
        self.stats = abi.RoundStats()
        self.num_jobs = inp.num_jobs

    ARRAYS = ("job_state", "job_node", "job_scheduled_at_priority", "job_preempted_at_priority", "job_method",
              "job_reason", "job_seq", "node_alloc", "queue_allocated", "queue_allocated_by_pc", "queue_fair_share",
              "scheduled_resources", "evicted_resources", "job_excluded_nodes")
    SCALARS = ("num_scheduled_jobs", "num_scheduled_gangs", "num_evicted_jobs", "termination_reason",
               "num_result_scheduled", "num_result_preempted")

    def diff(self, other: "RoundResult") -> List[str]:
        """Bit-exact comparison of every output; returns a list of human-readable mismatches."""
        bad = []
        names = self.ARRAYS + (("job_seq_first_pass", "job_reason_first_pass") if self.first_pass and other.first_pass else ())
        for name in names:
            x, y = getattr(self, name), getattr(other, name)
            if x.dtype.kind == "f":
                eq = x.view(np.uint64) == y.view(np.uint64)
            else:
                eq = x == y
            if not eq.all():
                idx = np.argwhere(~eq)[:5].tolist()
                bad.append(f"{name}: {int((~eq).sum())} mismatches, first at {idx}: {x[tuple(idx[0])]} vs {y[tuple(idx[0])]}")
        for name in self.SCALARS:
            if getattr(self.out, name) != getattr(other.out, name):
                bad.append(f"{name}: {getattr(self.out, name)} vs {getattr(other.out, name)}")
        return bad


def node_preemptibility_stats(b: "RoundInputBuilder", res: "RoundResult") -> List[Tuple[str, bool, str]]:
    """EvictorResult.NodePreemptiblityStats of the round's first evictor (NewNodeEvictor with the fair-share filter,
    preempting_queue_scheduler.go:93-136; eviction.go:197-273): per node, in node-id order, (node id, can every job on
    it be preempted, the sorted reasons).  A report for operators (scheduling reports, cycle metrics), derived on the
    host from the round's inputs and the fair shares the device computed — the evictions themselves happen on the
    device (`k_evict_balance`) and `evictable_jobs` below is checked against them in the tests."""
    reasons, _ = _eviction_reasons(b, res)
    inp = b.input
    by_node: Dict[int, List[int]] = {}
    for j, n in enumerate(b.job_node[: len(b.jobs)]):
        if n != abi.NONE:
            by_node.setdefault(int(n), []).append(j)
    out = []
    for i in sorted(range(len(b.nodes)), key=lambda i: b.nodes[i].id):
        node = b.nodes[i]
        jobs = by_node.get(i, [])
        if not jobs:  # nodeFilter: "node_empty" (eviction.go:88-95, 200-210)
            rs = {"node_empty"} | ({"node_unschedulable"} if node.unschedulable else set())
            out.append((node.id, not node.unschedulable, ",".join(sorted(rs))))
            continue
        rs = {reasons[j] for j in jobs if reasons[j]}
        if node.unschedulable:
            rs.add("node_unschedulable")
        out.append((node.id, not rs, ",".join(sorted(rs)) if rs else "all_jobs_preemptible"))
    return out


def evictable_jobs(b: "RoundInputBuilder", res: "RoundResult") -> np.ndarray:
    """The running jobs the first evictor's job filter lets through (bool[J])."""
    return _eviction_reasons(b, res)[1]


def _eviction_reasons(b: "RoundInputBuilder", res: "RoundResult"):
    cfg, inp = b.cfg, b.input
    J, D = len(b.jobs), b.factory.D
    total = b.total_resources.astype(np.float64)
    mult = np.array([inp.drf_multipliers[d] for d in range(D)])
    # qctx.GetAllocation() = Allocated + ShortJobPenalty at the start of the round
    alloc = b.qa.sum(axis=1).astype(np.float64) + b.qp.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        frac_d = np.where(total != 0, alloc / np.where(total != 0, total, 1.0), 0.0) * mult
        actual = np.maximum(frac_d.max(axis=1), 0.0) if D else np.zeros(len(b.queues))
        fs = res.queue_fair_share
        fair = fs[:, 2] if cfg.protect_uncapped_adjusted_fair_share else np.maximum(fs[:, 1], fs[:, 0])
        fraction = actual / fair[: len(actual)]
    reasons: List[str] = [""] * J
    evict = np.zeros(J, dtype=bool)
    for j, spec in enumerate(b.jobs):
        if b.job_node[j] == abi.NONE:
            continue
        q = b.job_queue[j]
        if q == abi.NONE:
            reasons[j] = "invalid_queue"
        elif not cfg.priority_classes[spec.priority_class].preemptible:
            reasons[j] = "job_not_preemptible"
        elif fraction[q] <= cfg.protected_fraction_of_fair_share:  # (NaN compares false: evicted, like the reference)
            reasons[j] = "below_protected_fair_share"
        else:
            evict[j] = True
    return reasons, evict


# constraints.go:17-55 + gang_scheduler.go:168-216: the strings behind ARMADA_REASON_*
REASON_TEXT = {
    abi.REASON_MAX_RESOURCES_SCHEDULED: "maximum resources scheduled",
    abi.REASON_MAX_RESOURCES_PER_QUEUE: "resource limit exceeded",
    abi.REASON_GLOBAL_RATE_LIMIT: "global scheduling rate limit exceeded",
    abi.REASON_QUEUE_RATE_LIMIT: "queue scheduling rate limit exceeded",
    abi.REASON_QUEUE_CORDONED: "queue cordoned",
    abi.REASON_GLOBAL_RATE_LIMIT_GANG: "gang would exceed global scheduling rate limit",
    abi.REASON_QUEUE_RATE_LIMIT_GANG: "gang would exceed queue scheduling rate limit",
    abi.REASON_GANG_EXCEEDS_GLOBAL_BURST: "gang cardinality too large: exceeds global max burst size",
    abi.REASON_GANG_EXCEEDS_QUEUE_BURST: "gang cardinality too large: exceeds queue max burst size",
    abi.REASON_GANG_DOES_NOT_FIT: "unable to schedule gang since minimum cardinality not met",
    abi.REASON_JOB_DOES_NOT_FIT: "job does not fit on any node",
    abi.REASON_UNIFORMITY_LABEL_NOT_INDEXED: "uniformity label is not indexed",
    abi.REASON_NO_NODES_WITH_UNIFORMITY_LABEL: "no nodes with uniformity label",
    abi.REASON_GANG_FITS_NO_UNIFORMITY_VALUE: "at least one job in the gang does not fit on any node",
    abi.REASON_FLOATING_RESOURCES: "floating resource limit",
}


@dataclass
class QueueStats:
    """scheduling.QueueStats (result.go:15-28) without the wall-clock `Time`."""
    gangs_considered: int = 0
    jobs_considered: int = 0
    gangs_scheduled: int = 0
    first_gang_considered_sample_job_id: str = ""
    first_gang_considered_result: str = ""
    first_gang_considered_queue_position: int = 0
    last_gang_scheduled_sample_job_id: str = ""
    last_gang_scheduled_queue_position: int = 0
    last_gang_scheduled_queue_cost: float = 0.0
    last_gang_scheduled_resources: Optional[np.ndarray] = None
    last_gang_scheduled_queue_resources: Optional[np.ndarray] = None


def queue_stats(b: "RoundInputBuilder", res: "RoundResult") -> Dict[str, QueueStats]:
    """SchedulingResult.AdditionalSchedulingInfo.StatsPerQueue: the statistics QueueScheduler.Schedule keeps per queue
    (queue_scheduler.go:190-235) for the FIRST schedule pass of the round (preempting_queue_scheduler.go:270-272), from
    the device's first-pass view (`job_seq_first_pass`, `job_reason_first_pass`).  A gang = the jobs of one loop
    iteration; its sample job = gctx.JobIds()[0], the first member in queue order; loopNumber = iteration − 1.  The
    queue's allocation and cost at its last scheduled gang are REPLAYED from the start-of-pass allocation (the
    snapshot's minus what the first evictor took) plus the first-pass successes up to that iteration."""
    if not res.first_pass:
        raise ValueError("queue_stats needs a RoundResult created with first_pass=True")
    J, f, cfg = len(b.jobs), b.factory, b.cfg
    seq, why = res.job_seq_first_pass[:J], res.job_reason_first_pass[:J]
    req = b.class_request[b.job_class[:J]] if J else np.zeros((0, f.D), np.int64)
    pcprio = np.array([cfg.priority_classes[j.priority_class].priority for j in b.jobs], np.int64)
    running = b.job_node[:J] != abi.NONE

    def order_key(j):  # SchedulingOrderCompare (jobdb/comparison.go:49-107)
        return (0 if running[j] else 1, -pcprio[j], b.job_qp[j], b.job_art[j] if running[j] else 0, b.job_st[j], b.job_id_rank[j])

    evicted = evictable_jobs(b, res)
    total = b.total_resources.astype(np.float64)
    mult = np.array([b.input.drf_multipliers[d] for d in range(f.D)])
    out: Dict[str, QueueStats] = {}
    for qi, q in enumerate(b.queues):
        mine = np.nonzero((b.job_queue[:J] == qi) & (seq > 0))[0]
        if len(mine) == 0:
            continue
        st = QueueStats()
        gangs: Dict[int, List[int]] = {}
        for j in mine:
            gangs.setdefault(int(seq[j]), []).append(int(j))
        alloc = b.qa[qi].sum(axis=0) + b.qp[qi]  # qctx.GetAllocation(): Allocated + ShortJobPenalty
        for j in np.nonzero((b.job_queue[:J] == qi) & evicted)[0]:
            alloc = alloc - req[j]
        for s in sorted(gangs):
            members = sorted(gangs[s], key=order_key)
            ok = all(why[j] == 0 for j in members)
            st.gangs_considered += 1
            st.jobs_considered += len(members)
            if st.first_gang_considered_sample_job_id == "":
                st.first_gang_considered_sample_job_id = b.jobs[members[0]].id
                st.first_gang_considered_queue_position = s - 1
                st.first_gang_considered_result = "scheduled" if ok else REASON_TEXT.get(int(why[members[0]]), str(int(why[members[0]])))
            if ok:
                st.gangs_scheduled += 1
                gang_total = req[members].sum(axis=0)
                alloc = alloc + gang_total
                st.last_gang_scheduled_sample_job_id = b.jobs[members[0]].id
                st.last_gang_scheduled_queue_position = s - 1
                st.last_gang_scheduled_resources = gang_total
                st.last_gang_scheduled_queue_resources = alloc.copy()
                with np.errstate(divide="ignore", invalid="ignore"):
                    frac = np.where(total != 0, alloc / np.where(total != 0, total, 1.0), 0.0) * mult
                cost = max(float(frac.max()), 0.0) if f.D else 0.0
                st.last_gang_scheduled_queue_cost = cost / float(b.qw[qi])
        out[q.name] = st
    return out
