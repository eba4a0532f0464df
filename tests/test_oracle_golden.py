"""CPU: pin the oracle against the reference.  tests/golden/*.npz are outputs of the reference's own CUDA kernels
(oracle/make_golden.py on a B200); the C restatement must reproduce them on the regenerated seeded inputs."""
import os

import numpy as np
import pytest

import helpers
import oracle_run

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden")
CASES = helpers.GOLDEN_RUNS


@pytest.mark.parametrize("name,far", CASES)
def test_oracle_reproduces_reference_kernels(name, far):
    fn = os.path.join(GOLDEN, f"{name}{'_far' if far else ''}.npz")
    assert os.path.exists(fn), "golden fixture missing: run oracle/make_golden.py on a GPU box"
    ref = dict(np.load(fn))
    kfs = helpers.build_case(name, far=far)
    orc = oracle_run.run_oracle(kfs, np.float32)
    np.testing.assert_allclose(orc["sig"], ref["sig"], rtol=1e-6, err_msg="seeded inputs differ from the golden run")
    for lt in helpers.MG_LOSSES:
        assert f"mmg_{lt}_AtA" in ref and f"mmg_{lt}_err_only" in ref, "golden fixture predates the match-geometry factors"
    assert "lmg_AtA" in ref and "lmg_err_only" in ref
    for k, v in ref.items():
        if k in ("sig",):
            continue
        e = helpers.rel_err(np.asarray(orc[k]).reshape(-1), np.asarray(v).reshape(-1))
        assert e <= 2e-5, f"{name} far={far}: {k} rel err {e:.3e}"
        if k.endswith("_AtA"):  # and block by block, every variable scaled to unit diagonal
            helpers.assert_blocks_close(k[:-4], orc[k], orc[k[:-4] + "_Atb"], v, ref[k[:-4] + "_Atb"], helpers.CASES[name]["C"], 1e-4,
                                        f"{name} far={far} oracle vs reference kernels")


def test_f32_and_f64_oracles_agree():
    kfs = helpers.build_case("small_c8_f16")
    a, b = oracle_run.run_oracle(kfs, np.float32), oracle_run.run_oracle(kfs, np.float64)
    for k in a:
        if k not in ("sig", "cam_pyramid"):
            assert helpers.rel_err(a[k], b[k]) < 1e-5, k
