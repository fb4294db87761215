"""Oracle pinned against the reference's own round-level known-answer tables:
  TestPreemptingQueueScheduler (scheduling/preempting_queue_scheduler_test.go:90-2377)
  TestQueueScheduler           (scheduling/queue_scheduler_test.go:32-682)
transcribed mechanically into tests/golden/*.json by tests/golden/extract_go_tables.py."""
import pytest

import go_tables as gt
import oracle_lib

PQS = gt.load_cases("preempting_queue_scheduler")
QS = gt.load_cases("queue_scheduler")


@pytest.mark.parametrize("name", sorted(PQS.keys()))
def test_preempting_queue_scheduler(name):
    try:
        gt.run_pqs_case(PQS[name], oracle_lib.round_schedule)
    except gt.UnsupportedCase as e:
        pytest.skip(f"not modelled: {e}")


@pytest.mark.parametrize("name", sorted(QS.keys()))
def test_queue_scheduler(name):
    try:
        gt.run_queue_scheduler_case(QS[name], oracle_lib.round_schedule)
    except gt.UnsupportedCase as e:
        pytest.skip(f"not modelled: {e}")
