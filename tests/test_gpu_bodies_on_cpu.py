"""Runs the BODIES of the golden-vector GPU tests of tests/test_gpu_widen.py on the CPU, with the device name patched to
"cpu" and the CUDA ops replaced by the oracle-backed stand-ins of tests/test_host_widen.py.  This does not test a kernel;
it makes sure, in the no-GPU suite, that the GPU tests themselves (fixtures, argument order, shapes, tolerances against the
golden vectors) are sound before they are spent on a GPU box."""
import pytest

import test_gpu_widen as G
from test_host_widen import cpu_ops  # noqa: F401  (fixture)

BODIES = ["test_edge_time_encoding_golden", "test_edge_forward_with_times_golden", "test_edge_forward_noisy_golden",
          "test_downprompt_node_golden", "test_downprompt_graph_golden", "test_library_build_golden",
          "test_graph_forward_golden", "test_inverse_sampling_golden", "test_fewshot_noisy_branch_shapes"]


@pytest.mark.parametrize("name", BODIES)
def test_gpu_test_body_runs_on_cpu(name, golden, cpu_ops, monkeypatch):  # noqa: F811
    monkeypatch.setattr(G, "DEV", "cpu")
    getattr(G, name)(golden)


@pytest.mark.parametrize("graph_level", [False, True])
def test_fewshot_forward_body_runs_on_cpu(graph_level, golden, cpu_ops, monkeypatch):  # noqa: F811
    monkeypatch.setattr(G, "DEV", "cpu")
    G.test_fewshot_forward_golden(golden, graph_level)


def test_ragged_readout_body_runs_on_cpu(cpu_ops, monkeypatch):  # noqa: F811
    monkeypatch.setattr(G, "DEV", "cpu")
    G.test_split_and_batchify_ragged()


def test_fewshot_helpers_body_runs_on_cpu(cpu_ops, monkeypatch):  # noqa: F811
    monkeypatch.setattr(G, "DEV", "cpu")
    G.test_fewshot_helpers_vs_oracle()
