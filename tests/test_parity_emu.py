"""Kernel-logic parity WITHOUT a GPU: the same CUDA source (armada_b200/csrc/*.cu|inc) compiled
by g++ against tools/simt_emu/cuda_emu.h (deterministic SIMT emulator, test tooling only) and
compared bit-for-bit with the CPU oracle.  The real parity tests are tests/test_parity_gpu.py
(pytest -m gpu, through the nvcc-built product library); this module exists so that control-flow /
data-structure bugs in the kernels are caught on a CPU-only box.  EMU_ORDER=reverse runs the
lanes of every warp in the opposite order (catches missing __syncwarp in either direction)."""
import os

import numpy as np

import pytest

import emu_lib
import go_tables as gt
import oracle_lib
from armada_b200 import abi, synth

_dev = None


def emu_round(inp):
    global _dev
    if _dev is None:
        _dev = emu_lib.emu_round()
    return _dev.schedule(inp)


def assert_parity(inp, label=""):
    want = oracle_lib.round_schedule(inp)
    got = emu_round(inp)
    bad = got.diff(want)
    assert not bad, f"{label}: emulated kernel != oracle:\n  " + "\n  ".join(bad)
    return got, want


@pytest.fixture(params=["forward", "reverse"])
def lane_order(request):
    old = os.environ.get("EMU_ORDER")
    os.environ["EMU_ORDER"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("EMU_ORDER", None)
    else:
        os.environ["EMU_ORDER"] = old


@pytest.mark.parametrize("seed", range(24))
def test_random_rounds(seed, lane_order):
    r = synth.random_round(
        seed,
        away=(seed % 4 == 1),
        round_limit=(seed % 6 == 3),
        queue_limits=(seed % 6 == 4),
        protected_fraction=0.5 if seed % 3 == 2 else 0.0,
        lookback=40 if seed % 5 == 1 else 0,
        n_nodes=40 + 13 * (seed % 7),
        n_jobs=300 + 50 * (seed % 5),
        n_running=80 + 20 * (seed % 4),
    )
    assert_parity(r.to_input(), r.name)


@pytest.mark.parametrize("seed", [100, 101])
def test_random_rounds_many_nodes(seed):
    r = synth.random_round(seed, n_nodes=1500 + 700 * (seed % 3), n_queues=9, n_jobs=2500, n_running=1200,
                           protected_fraction=0.5 if seed % 2 else 0.0)
    assert_parity(r.to_input(), r.name)


@pytest.mark.parametrize("n_queues", [40, 100, 128])
def test_many_queues(n_queues):
    """More queues than fit 64 batch items each (the batch windows shrink to 32 / the merge sort to
    a different power of two); priorities off so that most of the round runs in batch mode."""
    r = synth.random_round(300 + n_queues, n_nodes=300, n_queues=n_queues, n_jobs=2500, n_running=0, gangs=False, priorities=False)
    got, _ = assert_parity(r.to_input(), r.name)
    assert int(got.stats.phase_cycles[4]) > 0


@pytest.mark.parametrize("name,scale", [("C2", 0.02), ("C3", 0.004), ("C4", 0.006), ("C5", 0.004)])
def test_scaled_configs(name, scale):
    r = synth.scaled(name, scale)
    got, want = assert_parity(r.to_input(), f"{name}@{scale}")
    assert got.out.num_result_scheduled == want.out.num_result_scheduled


PQS = gt.load_cases("preempting_queue_scheduler")
QS = gt.load_cases("queue_scheduler")


def _emu_round_or_skip(inp):
    try:
        got = emu_round(inp)
    except abi.ArmadaError as e:
        if e.status == abi.E_UNSUPPORTED:
            raise gt.UnsupportedCase(str(e))
        raise
    want = oracle_lib.round_schedule(inp)
    bad = got.diff(want)
    assert not bad, "emulated kernel != oracle:\n  " + "\n  ".join(bad)
    return got


@pytest.mark.parametrize("name", sorted(PQS.keys()))
def test_reference_pqs_tables(name):
    try:
        gt.run_pqs_case(PQS[name], _emu_round_or_skip)
    except gt.UnsupportedCase as e:
        pytest.skip(f"outside device domain / not modelled: {e}")


@pytest.mark.parametrize("name", sorted(QS.keys()))
def test_reference_queue_scheduler_tables(name):
    try:
        gt.run_queue_scheduler_case(QS[name], _emu_round_or_skip)
    except gt.UnsupportedCase as e:
        pytest.skip(f"outside device domain / not modelled: {e}")


@pytest.mark.parametrize("n_nodes", [7, 401])
def test_runs_of_known_unschedulable_jobs(n_nodes):
    """synth.unfeasible_runs_round: skipped runs around the 128-record fast-forward step; the level
    scan that proves the miss covers node counts that are not a multiple of its stride."""
    r = synth.unfeasible_runs_round(n_nodes)
    got, want = assert_parity(r.to_input(), r.name)
    assert got.out.num_result_scheduled == want.out.num_result_scheduled == 11 + 40


@pytest.mark.parametrize("seed,indexed", [(400, [synth.CPU, synth.MEM]), (401, [synth.CPU]), (402, [synth.MEM, synth.GPU]),
                                          (403, [synth.CPU, synth.MEM])])
def test_partly_indexed_resources(seed, indexed, lane_order):
    """Not every resource is part of the best-fit key (nodedb indexedResources ⊂ resources): the
    key no longer carries the whole row, so the SWAR shortcuts are off and the assignment table
    keeps rows beside the keys (table_assign<false>, row-reading cursor refills)."""
    batchy = seed != 403
    r = synth.random_round(seed, n_nodes=120, n_queues=6, n_jobs=900, n_running=0 if batchy else 200, gangs=not batchy, priorities=not batchy)
    r.indexed = indexed
    got, _ = assert_parity(r.to_input(), f"{r.name} indexed={indexed}")
    if batchy:
        assert int(got.stats.phase_cycles[4]) > 0


@pytest.fixture(params=["k32", "k64"])
def compare_key(request):
    """The assignment loop compares 32-bit compact keys when the resource fields fit 26 bits
    (ARMADA_NO_K32 forces the general 64-bit form)."""
    if request.param == "k64":
        os.environ["ARMADA_NO_K32"] = "1"
    yield request.param
    os.environ.pop("ARMADA_NO_K32", None)


@pytest.mark.parametrize("wq,seed", [(8, 500), (8, 501), (16, 502), (8, 503)])
def test_long_batch_pipelines(wq, seed, lane_order, compare_key):
    """Small batches (ARMADA_BT_WQ test knob: items per queue per batch) make one pipeline run span
    many batches: batch k+1 is produced from the speculative queue state while batch k is assigned,
    jobs that find no node cut a batch short (the batch built behind it is dropped), runs of
    known-unschedulable jobs sit between committed items."""
    os.environ["ARMADA_BT_WQ"] = str(wq)
    try:
        dev = emu_lib.emu_round()
        if seed == 503:
            r = synth.unfeasible_runs_round(7)
        else:
            r = synth.random_round(seed, n_nodes=24 + 10 * (seed % 3), n_queues=5 + seed % 4, n_jobs=1400, n_running=0, gangs=seed == 502,
                                   priorities=False, round_limit=seed == 501)
        inp = r.to_input()
        want = oracle_lib.round_schedule(inp)
        got = dev.schedule(inp)
        bad = got.diff(want)
        assert not bad, f"{r.name}: emulated kernel != oracle:\n  " + "\n  ".join(bad)
        if seed != 503:
            assert int(got.stats.batch_cycles[6]) >= {500: 8, 501: 2, 502: 3}[seed], "expected several batches per pipeline run"
        dev.close()
    finally:
        os.environ.pop("ARMADA_BT_WQ", None)


@pytest.mark.parametrize("nodes,queues,jobs,wq,seed", [(60, 4, 1500, 0, 1), (120, 8, 3000, 8, 2), (250, 6, 5000, 16, 3), (40, 3, 900, 0, 4)])
def test_gangs_as_batch_items(nodes, queues, jobs, wq, seed, lane_order, compare_key):
    """Simple gangs (complete, one class, contiguous in their queue — what the C4 generator makes) are
    ordered by the batch pipeline as ONE item and placed member by member by the assignment loop, all or
    nothing: the clusters here fill up, so gangs fail in the middle (roll-back of the table, of the
    touched bits and of the window a refill replaced) and the general loop fails them the reference's way."""
    if wq:
        os.environ["ARMADA_BT_WQ"] = str(wq)
    try:
        dev = emu_lib.emu_round()
        r = synth.config_c4(nodes, queues, jobs, seed=synth.SEED + seed)
        inp = r.to_input()
        want = oracle_lib.round_schedule(inp)
        got = dev.schedule(inp)
        bad = got.diff(want)
        assert not bad, f"C4 {nodes}x{jobs}: emulated kernel != oracle:\n  " + "\n  ".join(bad)
        # gang members were placed in batch mode: more placements than iterations there
        assert int(got.stats.placements) > int(got.stats.loop_iterations) - int(np.count_nonzero(np.asarray(want.job_state) == 4))
        assert int(got.stats.phase_cycles[4]) > 0
        dev.close()
    finally:
        os.environ.pop("ARMADA_BT_WQ", None)


@pytest.mark.parametrize("seed", range(10))
def test_exact_mode_unaligned_rounds(seed, lane_order):
    """Inputs outside the fast domain run in exact mode: the reference's default index resolutions
    (cpu 100m, memory 100Mi) with 250m / 4Gi-style requests, node sizes that are not multiples of
    them, allocatable < total, a NodeFactory index order that differs from the node-id order, classes
    that match several node types.  Every probe is the literal ordered walk of nodeiteration.go."""
    r = synth.random_round(seed, away=(seed % 4 == 1), n_nodes=40 + 13 * (seed % 7), n_jobs=300 + 50 * (seed % 5),
                           n_running=80 + 20 * (seed % 4), protected_fraction=0.5 if seed % 3 == 2 else 0.0,
                           round_limit=(seed % 6 == 3), queue_limits=(seed % 6 == 4), lookback=40 if seed % 5 == 1 else 0, unaligned=True)
    assert_parity(r.to_input(), r.name)


def test_exact_mode_resolution_rounding_blocks_a_feasible_node():
    """gang_scheduler_test.go:244-262 on the device path: the fourth job fits a node but the rounded
    index key hides that node from the iterator."""
    got, want = assert_parity(synth.rounding_round().to_input(), "rounding")
    assert got.out.num_result_scheduled == want.out.num_result_scheduled == 3
    assert int((got.job_state == abi.JOB_FAILED).sum()) == 1


@pytest.mark.parametrize("seed", [0, 5])
def test_exact_mode_forced_on_aligned_rounds(seed):
    """ARMADA_FORCE_EXACT: the literal walk on inputs the fast path also accepts gives the same round."""
    os.environ["ARMADA_FORCE_EXACT"] = "1"
    try:
        dev = emu_lib.emu_round()
        r = synth.random_round(seed, away=(seed % 4 == 1), n_nodes=60, n_jobs=350, n_running=90, protected_fraction=0.5 if seed else 0.0)
        inp = r.to_input()
        assert not dev.schedule(inp).diff(oracle_lib.round_schedule(inp))
        dev.close()
    finally:
        os.environ.pop("ARMADA_FORCE_EXACT", None)


def test_more_classes_than_the_shared_memory_table_holds():
    r = synth.many_classes_round()
    got, _ = assert_parity(r.to_input(), r.name)
    assert got.out.num_result_scheduled > 0


def test_time_budget_aborts_the_round_and_leaves_the_snapshot_runnable():
    """armada_round_run_deadline: a budget that cannot be met returns ARMADA_E_DEADLINE (the reference's
    cancelled context, scheduling_algo.go:115-118), download is refused, and the same handle still
    schedules the uploaded snapshot afterwards — bit-exact.  (Emulated kernel build: no GPU needed.)"""
    r = synth.random_round(3, n_nodes=50, n_jobs=300, n_running=60)
    inp = r.to_input()
    dev = emu_lib.emu_round()
    dev.upload(inp)
    with pytest.raises(abi.ArmadaError) as ei:
        dev.run(budget_ns=1)
    assert ei.value.status == abi.E_DEADLINE
    with pytest.raises(abi.ArmadaError) as ei:
        dev.download()
    assert ei.value.status == abi.E_STATE
    dev.run(budget_ns=60_000_000_000)
    assert not dev.download().diff(oracle_lib.round_schedule(inp))
    dev.close()


def _dry_run_case(r, gangs_as_jobs, lib):
    """gangs_as_jobs: lists of job indices; the product takes the jobs' classes."""
    from armada_b200.scheduler import DeviceNodeDb
    inp = r.to_input()
    jc = np.asarray(r.job_class).astype(np.int64)
    odb = oracle_lib.OracleNodeDb(inp)
    want_ok, want_nodes = [], []
    for g in gangs_as_jobs:
        ok, nodes = odb.dry_run(g)
        want_ok.append(ok)
        want_nodes.append(nodes)
    with DeviceNodeDb(inp, 0, lib=lib) as db:
        got_ok, got_nodes = db.schedule_many([[int(jc[j]) for j in g] for g in gangs_as_jobs])
        # armada_nodedb_select_nodes: the single-job form gives the same nodes as gangs of one
        singles = [g for g in gangs_as_jobs if len(g) == 1]
        sel = db.select_nodes([int(jc[g[0]]) for g in singles])
        for g, n in zip(singles, sel):
            ok, nodes = odb.dry_run(g)
            assert (n != abi.NONE) == ok and (not ok or n == nodes[0])
    assert list(got_ok) == want_ok
    for g, (a, b) in enumerate(zip(got_nodes, want_nodes)):
        assert (a == b).all(), f"gang {g}: {a} vs {b}"
    return want_ok


@pytest.mark.parametrize("seed,unaligned", [(0, False), (1, True), (2, True), (3, False)])
def test_dry_run_nodedb_matches_the_oracle(seed, unaligned):
    """armada_nodedb_schedule_many (SubmitChecker's ScheduleManyWithTxn + Abort on an empty cluster,
    submitcheck.go:302-422): single jobs and gangs of every class, including gangs too big for the
    cluster, against the oracle's NodeDb — same verdicts, same nodes."""
    r = synth.random_round(seed, n_nodes=30 + 7 * seed, n_jobs=200, n_running=0, gangs=False, unaligned=unaligned, away=(seed == 1))
    rng = np.random.default_rng(seed)
    J = len(np.asarray(r.job_class))
    gangs = [[int(j)] for j in rng.choice(J, 40, replace=False)]
    for _ in range(12):
        size = int(rng.integers(2, 150))
        j0 = int(rng.integers(0, J))
        cls = np.asarray(r.job_class)[j0]
        same = np.nonzero(np.asarray(r.job_class) == cls)[0]
        gangs.append([int(same[i % len(same)]) for i in range(size)] if len(same) >= 1 else [j0])
    # the oracle's jobs must be distinct inside one gang
    gangs = [list(dict.fromkeys(g)) for g in gangs]
    oks = _dry_run_case(r, gangs, emu_lib.load())
    assert any(oks)


def test_dry_run_nodedb_resolution_rounding():
    """The rounded index key can hide a node from the iterator even on an empty cluster: 20 cpu on a
    32-cpu node with a 17-cpu index resolution (rounded 17 < 20) is unschedulable in the reference."""
    r = synth.rounding_round()
    r.class_request = np.stack([synth.rl(16, 128), synth.rl(20, 128)])
    r.class_pc = np.zeros(2)
    r.class_static_row = np.zeros(2)
    r.job_class = np.array([0, 0, 1, 1])
    oks = _dry_run_case(r, [[0], [2], [0, 1], [3]], emu_lib.load())
    assert oks == [True, False, True, False]


@pytest.mark.parametrize("seed", [2, 4, 10])
def test_snapshot_construction_on_the_device(seed):
    """NULL queue_allocated_by_pc / queue_constrained_demand: the library derives the queue accounting
    from the job arrays (calculateJobSchedulingInfo + constructSchedulingContext,
    scheduling_algo.go:522-632,664-676) — on the device in the product (k_snapshot_*), restated in the
    oracle; without per-queue limits that equals what the host-side generator passes explicitly."""
    limits = seed == 4
    r = synth.random_round(seed, n_nodes=50, n_jobs=350, n_running=100, protected_fraction=0.5, queue_limits=limits)
    inp = r.to_input()
    explicit = oracle_lib.round_schedule(inp)
    inp.queue_allocated_by_pc = None
    inp.queue_constrained_demand = None
    want = oracle_lib.round_schedule(inp)
    got = emu_round(inp)
    assert not got.diff(want)
    if not limits:
        assert not want.diff(explicit)


def _excluded_nodes_properties(inp, want):
    """queue_scheduler_test.go:656-676: for a single job that could not be scheduled the excluded nodes
    add up to the number of nodes; jobs that were never attempted (or did not fail) report nothing."""
    ex = np.asarray(want.job_excluded_nodes)
    st = np.asarray(want.job_state)
    gang = np.ctypeslib.as_array(inp.job_gang, (inp.num_jobs,))
    tot = ex.sum(axis=1)
    assert (tot[(st != abi.JOB_FAILED) | (gang != abi.NONE)] == 0).all()
    attempted = tot > 0
    assert (tot[attempted] == inp.num_nodes).all()
    return int(attempted.sum())


@pytest.mark.parametrize("seed", range(8))
def test_excluded_nodes_by_reason_kind(seed, lane_order):
    """collect_excluded_nodes: PodSchedulingContext.NumExcludedNodesByReason of the jobs that fail, by
    reason kind (node type / taints+labels / resources on a reached node / never reached), identical to
    the oracle's restatement of nodedb.go:445-480,605-640,786-797,1102-1117."""
    r = synth.random_round(700 + seed, n_nodes=40 + 7 * seed, n_queues=4, n_jobs=600, n_running=100 if seed % 2 else 0, gangs=seed % 3 == 0,
                           priorities=seed % 2 == 1, unaligned=seed >= 6)
    inp = r.to_input()
    inp.collect_excluded_nodes = 1
    want = oracle_lib.round_schedule(inp)
    got = emu_round(inp)
    assert not got.diff(want)
    assert _excluded_nodes_properties(inp, want) > 0 or seed in (0,)


def test_excluded_nodes_on_the_reference_tables():
    """The same on the reference's QueueScheduler table (the test that states the property)."""
    seen = []

    def schedule(inp):
        inp.collect_excluded_nodes = 1
        got = _emu_round_or_skip(inp)  # (compares every array, job_excluded_nodes included, with the oracle)
        seen.append(_excluded_nodes_properties(inp, got))
        return got

    for name in sorted(QS.keys()):
        try:
            gt.run_queue_scheduler_case(QS[name], schedule)
        except gt.UnsupportedCase:
            continue
    assert sum(seen) > 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_excluded_nodes_static_and_resource_kinds(seed, lane_order):
    """Static classes finer than node types (a taint / label the node type index does not carry): nodes of a
    matching node type that the walk reaches and StaticJobRequirementsMet rejects count under
    ARMADA_EXCL_STATIC; with a resource that is not indexed, reached nodes can also fail on resources."""
    r = synth.random_round(740 + seed, n_nodes=60, n_queues=3, n_jobs=500, n_running=0, gangs=False, priorities=False)
    if seed == 2:
        r.indexed = [synth.CPU, synth.MEM]  # gpu is not in the index: a reached node can still lack gpus
    # half of the nodes of type 0 get a new static class that row 0 rejects (type_match still accepts their type)
    sc = np.asarray(r.node_static_class).astype(np.uint32).copy()
    typ = np.asarray(r.node_type).astype(np.uint32)
    new = r.num_static_classes
    pick = np.nonzero(typ == 0)[0][::2]
    sc[pick] = new
    sm = np.asarray(r.static_match).astype(np.uint32)
    rows = sm.shape[0]
    words = (new + 1 + 31) // 32
    sm2 = np.zeros((rows, words), np.uint32)
    for row in range(rows):
        for c in range(new):
            if (sm[row, c >> 5] >> (c & 31)) & 1:
                sm2[row, c >> 5] |= np.uint32(1 << (c & 31))
        if row != 0 and (sm[row, 0] & 1):  # the new class behaves like class 0, except for row 0
            sm2[row, new >> 5] |= np.uint32(1 << (new & 31))
    r.node_static_class, r.static_match, r.num_static_classes = sc, sm2, new + 1
    inp = r.to_input()
    inp.collect_excluded_nodes = 1
    want = oracle_lib.round_schedule(inp)
    got = emu_round(inp)
    assert not got.diff(want)
    assert _excluded_nodes_properties(inp, want) > 0
    ex = np.asarray(want.job_excluded_nodes)
    assert ex[:, abi.EXCL_STATIC].sum() > 0


# ---- gang node uniformity + floating resources (gang_scheduler.go:143,154-223) ----------------------
import gang_cases  # noqa: E402


@pytest.mark.parametrize("name", sorted(gang_cases.GANG.keys()))
def test_reference_gang_scheduler_table(name):
    """TestGangScheduler (gang_scheduler_test.go:33-760) through the kernel: node-uniformity search, floating
    resources, round / queue limits, the resolution-rounding case."""
    b, tc, gangs = gang_cases.gang_case_round(name)
    got, _ = assert_parity(b.input, name)
    gang_cases.check_gang_case(b, tc, gangs, got)


@pytest.mark.parametrize("seed", range(8))
def test_uniformity_and_floating_rounds(seed, lane_order):
    got, want = assert_parity(gang_cases.uniformity_round(seed).input, f"uniformity round {seed}")
    if seed == 1:  # the generator reaches every new outcome
        reasons = set(int(x) for x in want.job_reason)
        assert {abi.REASON_UNIFORMITY_LABEL_NOT_INDEXED, abi.REASON_NO_NODES_WITH_UNIFORMITY_LABEL, abi.REASON_GANG_FITS_NO_UNIFORMITY_VALUE,
                abi.REASON_FLOATING_RESOURCES} <= reasons


@pytest.mark.parametrize("seed", [20, 21, 22])
def test_uniformity_rounds_in_exact_mode(seed):
    assert_parity(gang_cases.uniformity_round(seed, unaligned=True).input, f"unaligned uniformity round {seed}")


def test_uniformity_rounds_without_floating_resources_keep_the_batch_pipeline():
    got, _ = assert_parity(gang_cases.uniformity_round(30, n_nodes=80, n_jobs=600, floating=False).input, "uniformity, no floating")


import order_cases  # noqa: E402


@pytest.mark.parametrize("name", sorted(order_cases.CASES))
def test_job_priority_comparer(name):
    b, expected = order_cases.comparison_round(name)
    got, _ = assert_parity(b.input, name)
    order_cases.check_order(b, expected, got)


def test_presorted_job_order_fast_path_equals_the_general_path(monkeypatch):
    """armada_round_upload ranks the jobs of a queue without building sort keys when every queue already arrives in
    SchedulingOrderCompare order (the BASELINE configs do); ARMADA_NO_PRESORTED forces the bucket sort."""
    for name, scale in (("C3", 0.004), ("C5", 0.004), ("C4", 0.006)):
        inp = synth.scaled(name, scale).to_input()
        fast, want = assert_parity(inp, f"{name} presorted")
        monkeypatch.setenv("ARMADA_NO_PRESORTED", "1")
        general, _ = assert_parity(inp, f"{name} general order path")
        monkeypatch.delenv("ARMADA_NO_PRESORTED")
        assert not fast.diff(general)


def test_job_order_over_several_host_slices():
    """70 000 jobs: the upload ranks them on 16 host threads.  (1) queues that arrive in order: the one-pass ranking,
    checked inside every slice and across the slices of a queue; (2) the same round with two jobs of one queue
    exchanged ACROSS a slice boundary (only the cross-slice check can see it) and (3) inside a slice: the bucket sort."""
    r = synth.scaled("C2", 0.7)
    assert len(r.job_queue) >= 65536
    assert_parity(r.to_input(), "presorted, 16 slices")
    per_slice = len(r.job_queue) // 16
    for name, lo, hi in (("across slices", per_slice - 40, per_slice + 40), ("inside a slice", 100, 180)):
        r = synth.scaled("C2", 0.7)
        q = r.job_queue[lo]
        a = next(j for j in range(lo, hi) if r.job_queue[j] == q and j < per_slice) if name == "across slices" else lo
        b = next(j for j in range(hi, lo, -1) if r.job_queue[j] == q and (j >= per_slice or name != "across slices"))
        assert a < b and r.job_queue[a] == r.job_queue[b]
        r.job_submit_time[a], r.job_submit_time[b] = r.job_submit_time[b], r.job_submit_time[a]
        got, want = assert_parity(r.to_input(), name)
        assert want.job_seq[b] < want.job_seq[a] or want.job_seq[a] == 0  # b now comes first in its queue


@pytest.mark.parametrize("seed", [50, 51, 52, 53])
def test_uniformity_rounds_with_large_and_running_gangs(seed):
    """Gangs of up to 40 members (more than one warp load of members per attempt) and running gangs with a uniformity
    label that are evicted and re-scheduled (pinned members inside the per-value attempts)."""
    b = gang_cases.uniformity_round(seed, n_nodes=60, n_zones=4, n_jobs=500, floating=(seed % 2 == 0), max_gang=40, running_gangs=6,
                                    protected_fraction=0.0 if seed >= 52 else 0.5)
    got, want = assert_parity(b.input, f"large / running uniformity gangs {seed}")
    if seed >= 52:
        assert int(want.stats.evicted_pass1) > 0  # the running gangs are evicted and go through the search pinned


def test_dry_run_nodedb_ignores_floating_resources_in_the_node_fit():
    """SubmitChecker's NodeDb with a floating resource in the factory: a job's floating request is no part of the node
    fit (KubernetesResourceRequirements) — oracle and device agree, and the job fits where its cpu / memory do."""
    from armada_b200.scheduler import DeviceNodeDb

    class _R:  # what _dry_run_case needs of a round
        def __init__(self, b):
            self.b, self.job_class = b, b.job_class

        def to_input(self):
            return self.b.input

    from armada_b200.model import RoundInputBuilder
    b0 = gang_cases.uniformity_round(3, n_nodes=12, n_jobs=80, floating=True)
    # (the SubmitChecker's NodeDb is the EMPTY cluster: no running jobs in the input)
    b = RoundInputBuilder(b0.cfg, b0.nodes, [j for j in b0.jobs if j.node is None], b0.queues)
    J = len(b.jobs)
    with_lic = [j for j in range(J) if "licences" in b.jobs[j].requests and b.jobs[j].node is None]
    assert with_lic
    gangs = [[j] for j in with_lic[:10]] + [[j] for j in range(J) if b.jobs[j].node is None][:10]
    oks = _dry_run_case(_R(b), gangs, emu_lib.load())
    assert all(oks[: len(with_lic[:10])])  # 1–16 cpu on 32-cpu nodes of an empty cluster: every one fits
