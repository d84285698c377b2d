"""The driver's entry points: __graft_entry__.smoke() must run as-is on cuda:0."""
import pytest

pytestmark = pytest.mark.gpu


def test_graft_entry_smoke(capsys):
    import __graft_entry__ as g
    g.smoke()
    assert "smoke:" in capsys.readouterr().out
