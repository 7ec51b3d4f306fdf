"""ggs_b200.numa: the cpulist parser and the never-raises contract (host logic, no GPU)."""
import os

from ggs_b200 import numa


def test_parse_cpulist():
    assert numa._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert numa._parse_cpulist("") == []
    assert numa._parse_cpulist("5") == [5]


def test_bind_never_raises_and_keeps_affinity_without_a_gpu(monkeypatch):
    before = os.sched_getaffinity(0)
    monkeypatch.setattr(numa, "device_numa_node", lambda i: None)
    info = numa.bind_to_device(0)
    assert info["bound"] is False
    assert os.sched_getaffinity(0) == before


def test_bind_to_the_local_node(monkeypatch, tmp_path):
    """Two fake nodes: the process ends up on the CPUs of the GPU's node that it was allowed before."""
    before = os.sched_getaffinity(0)
    cpus = sorted(before)
    if len(cpus) < 2:
        return
    half = cpus[: len(cpus) // 2]
    real_open, real_listdir = open, os.listdir

    def fake_listdir(path):
        return ["node0", "node1", "online"] if path == "/sys/devices/system/node" else real_listdir(path)

    def fake_open(path, *a, **k):
        if path == "/sys/devices/system/node/node1/cpulist":
            f = tmp_path / "cpulist"
            f.write_text(",".join(str(c) for c in half) + "\n")
            return real_open(f, *a, **k)
        return real_open(path, *a, **k)

    monkeypatch.setattr(numa, "device_numa_node", lambda i: 1)
    monkeypatch.setattr(os, "listdir", fake_listdir)
    monkeypatch.setattr("builtins.open", fake_open)
    try:
        info = numa.bind_to_device(0)
        assert info == {"node": 1, "cpus": len(half), "bound": True}
        assert os.sched_getaffinity(0) == set(half)
    finally:
        os.sched_setaffinity(0, before)
