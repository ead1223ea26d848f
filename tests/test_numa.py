"""Host placement helper (merlot_reserve_b200.numa): cpulist parsing; binding is a no-op without a GPU."""
from merlot_reserve_b200 import numa


def test_parse_cpulist():
    assert numa._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert numa._parse_cpulist("") == []


def test_bind_without_gpu_is_harmless():
    import os
    before = os.sched_getaffinity(0)
    numa.bind_to_gpu(0)          # no CUDA device here: must leave the affinity alone or shrink it to a valid subset
    assert os.sched_getaffinity(0) <= before and len(os.sched_getaffinity(0)) > 0
    os.sched_setaffinity(0, before)
