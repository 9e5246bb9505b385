"""The driver's smoke entry (forward against the oracle, then one training step incl. the LayerNorm against float64
autograd) must pass as a test as well."""
import pytest

pytestmark = pytest.mark.gpu


def test_graft_entry_smoke():
    import __graft_entry__ as g
    g.smoke()
