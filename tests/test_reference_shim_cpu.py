"""CPU: the reference-side binding (oracle/ref_b200_shim.cpp: the unmodified reference's PhyloTree with its kernel pointers bound to
the C ABI) with the C ABI served by the CPU oracle (oracle/b200lk_over_oracle.c; oracle/_ref/libiqref_over_oracle.so).  What is
checked here is the binding's own logic — slot numbers from partial_lh pointers (central_partial_lh and nni_partial_lh), marshalling
of the reference's traversal_info, theta_computed, the optimizeOneBranch hook, the +ASC columns — under the reference's own
optimizeAllBranches / getBestNNIForBran / optimizeTreeLengthScaling, against the same reference over its AVX/FMA kernel.  The same
test functions run against the CUDA library in tests/test_gpu_reference_shim.py."""
import os

import pytest

from oracle import ref_engine

OVER = os.path.join(os.path.dirname(ref_engine.LIB), "libiqref_over_oracle.so")
pytestmark = pytest.mark.skipif(not (ref_engine.available() and os.path.exists(OVER)), reason="oracle/_ref libraries not built")


@pytest.fixture()
def over_oracle(monkeypatch):
    monkeypatch.setattr(ref_engine, "LIB_B200", OVER)
    monkeypatch.setattr(ref_engine, "_lib_b200", None)
    yield
    ref_engine._lib_b200 = None


@pytest.mark.parametrize("kind", ["config1", "dna_inv", "aa_c4", "bin"])
def test_binding_under_the_references_control_flow(over_oracle, kind):
    import test_gpu_reference_shim as m
    m.test_reference_control_flow_over_the_cuda_kernels(kind)


def test_binding_with_the_fused_newton_hook_and_asc(over_oracle):
    import test_gpu_reference_shim as m
    m.test_reference_with_the_fused_newton_hook("config1")
    m.test_reference_with_asc_over_the_cuda_kernels(1)
    m.test_reference_with_asc_over_the_cuda_kernels(2)


def test_binding_on_a_rooted_tree_with_a_nonreversible_model(over_oracle):
    import test_gpu_reference_shim as m
    m.test_reference_rooted_tree_with_a_nonreversible_model_over_the_cuda_kernels(4, 4, False)
    m.test_reference_rooted_tree_with_a_nonreversible_model_over_the_cuda_kernels(20, 2, True)
