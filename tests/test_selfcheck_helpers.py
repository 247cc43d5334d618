"""CPU test of the helpers bench.py's self_check and the 60-channel GPU test rely on (tests/util.py): the loop-chain
replay and the sampled one-step parity must accept the oracle's own closed-loop trajectory, and must reject a
trajectory with a perturbed discriminator or a perturbed correlator sum."""
import numpy as np
import pytest

import bds_oracle as O
import util


def _oracle_planes(mode, n_epochs):
    s, sats, x, ch = util.record(mode, 2, 0.01 * (n_epochs + 2.5) if mode != "B2a" else 0.001 * (n_epochs + 2.5))
    tr, raw = util.oracle_track(mode, s, x, ch, n_epochs)
    names = ["absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L", "Pilot_I_P", "Pilot_I_E",
             "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_P", "Pilot_Q_L", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt",
             "remCodePhase", "remCarrPhase"]
    planes = {n: np.array([t[n] if n in t else np.zeros(n_epochs) for t in tr]) for n in names}
    return s, x, ch, planes, raw


@pytest.mark.parametrize("mode", ["WB", "NB"])
def test_replay_and_sampled_parity_accept_the_oracle_trajectory(mode):
    n = 5
    s, x, ch, planes, raw = _oracle_planes(mode, n)
    so = O.Settings(dict(s))
    mem = util.replay_loop_chain(mode, so, ch, planes, n)
    for c in range(len(ch)):
        g = {k: planes[k][c] for k in planes}
        w = util.one_step_parity_sampled(mode, so, lambda pos, m: x[pos:pos + m], ch[c], g, raw[c], mem[:, :, c], [n - 2, n - 1])
        assert w <= 1e-9
    bad = {k: v.copy() for k, v in planes.items()}
    bad["pllDiscr"][0, 2] += 1e-6
    with pytest.raises(AssertionError):
        util.replay_loop_chain(mode, so, ch, bad, n)
    raw2 = raw.copy()
    raw2[0, n - 1, 2] *= 1.0 + 3e-4          # I_P of the data family off by 3e-4
    g = {k: planes[k][0] for k in planes}
    with pytest.raises(AssertionError):
        util.one_step_parity_sampled(mode, so, lambda pos, m: x[pos:pos + m], ch[0], g, raw2[0], mem[:, :, 0], [n - 1])
