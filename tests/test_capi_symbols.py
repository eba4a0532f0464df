"""CPU: the C-ABI library loads and exports every symbol include/sage_ba.h declares (no compute calls)."""
import os
import re

import helpers
import sage_slam_b200 as sage


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(helpers.ROOT, "include", "sage_ba.h")).read()
    declared = set(re.findall(r"\b(sage_ba_[a-z_0-9]+)\s*\(", hdr)) - {"sage_ba_allreduce_fn"}
    lib = sage.capi.load()
    assert declared == set(sage.capi.SIGNATURES), declared ^ set(sage.capi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sage_ba_version().startswith(b"sage-ba-b200")


def test_context_creation_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        return
    import pytest

    with pytest.raises(sage.SageError):
        sage.Context(0)


def test_product_never_imports_the_oracle():
    """The product package must not import, link or execute anything under oracle/ (docstrings may mention it)."""
    pkg = os.path.join(helpers.ROOT, "sage-slam_b200")
    pat = re.compile(r"^\s*(import|from)\s+(oracle|build_ref|oracle_run)\b|libsage_oracle|sage_oracle\.c|CDLL\([^)]*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f


def test_lineariser_choice_is_a_process_wide_switch_that_needs_no_gpu():
    """sage_ba_set_geometric_tcgen05: query with a negative argument, set returns the previous value (include/sage_ba.h)."""
    lib = sage.capi.load()
    prev = lib.sage_ba_set_geometric_tcgen05(-1)
    assert prev in (0, 1)
    assert lib.sage_ba_set_geometric_tcgen05(0) == prev
    assert lib.sage_ba_set_geometric_tcgen05(-1) == 0
    assert lib.sage_ba_set_geometric_tcgen05(1) == 0
    assert lib.sage_ba_set_geometric_tcgen05(prev) == 1
    assert lib.sage_ba_set_geometric_tcgen05(-1) == prev
