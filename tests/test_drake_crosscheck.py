"""analytic models against Drake -- only where pydrake exists (skipped in this image)."""
import pytest

pytest.importorskip("pydrake")

from tools import drake_crosscheck  # noqa: E402


@pytest.mark.parametrize("name", sorted(drake_crosscheck.CASES))
def test_model_matches_drake_discrete_update(name):
    # the stock plants are integrated semi-implicitly by discrete MultibodyPlant; the analytic
    # models restate that map, so the gap should be at rounding level times conditioning
    w = drake_crosscheck.crosscheck(name, samples=20)
    assert w["step"] < 1e-8, w
    assert w["fx"] < 1e-6 and w["fu"] < 1e-6, w
