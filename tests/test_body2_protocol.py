"""The barrier protocol of the persistent body kernel (csrc/body2_umma.cuh), checked on the CPU with the discrete-event
model of tools/body2_protocol_sim.py: the protocol as shipped admits no wrong tile under random heavy-tailed latencies,
and the model is sensitive - the protocol as it was before the ring-barrier fix is caught with the
signature tools/soak2.py saw on hardware (issuer B's last tile of a pass reads a slot whose box has not landed)."""
import os, sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import body2_protocol_sim as sim

CTAS = [1, 19, 20, 70, 128, 147]


def test_shipped_protocol_has_no_violation():
    for mode in ("both", "two", "count"):
        assert sim.sweep(CTAS, 1, 4, mode, regimes=(4, 16)) == []


def test_build_options_are_deadlock_free():
    # FEN_B2_TURN = 0 (issuers run concurrently) and FEN_B2_ROTATE = 1 (tile shares rotate per pass)
    assert sim.sweep(CTAS[:3], 1, 4, "both", regimes=(4,), turn=False) == []
    assert sim.sweep(CTAS[:3], 1, 4, "both", regimes=(4,), rotate=True) == []
    assert sim.sweep(CTAS[:3], 1, 5, "both", regimes=(4,), se_self=False) == []     # FEN_B2_SE_SELF = 0: SE batches issued by issuer A


def test_model_catches_the_protocol_before_the_fix():
    bad = sim.sweep(CTAS, 1, 4, "nofix", regimes=(4, 16), turn=False, se_self=False)
    assert bad, "the model no longer reproduces the phase aliasing of the single-barrier ring"
    assert all("issuer 1" in b and "inflight=True" in b for b in bad), bad[:3]
