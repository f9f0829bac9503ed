"""The generic BeamSearch mirror (models/allennlp_beamsearch.py; reference allennlp_beamsearch.py:39-294) against the oracle's
restatement of the reference algorithm, with a synthetic Markov step function: normal runs, early stop (variable
`steps_taken`), the beam_size == 1 all-<end> early return (:130-136), the per_node_beam_size > V configuration error
(:119-124), the infinite-log-probability warning (:263-268), and random shapes (hypothesis).  Kernels emulated on CPU."""
import warnings

import pytest
import torch
from hypothesis import given, settings, strategies as st

from dlsg import ops
from dlsg.errors import ConfigurationError
from oracle import dlsg_oracle as O
from cpu_emul import CpuEmulBackend

END = 2


@pytest.fixture(autouse=True)
def emul_backend():
    old = ops._backend
    ops.set_backend(CpuEmulBackend())
    yield
    ops.set_backend(old)


def markov(V, seed, end_bias=0.0, valid=None):
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(V, V, generator=g) * 2.0
    table[:, END] += end_bias
    if valid is not None:                         # only `valid` tokens ever get probability mass
        mask = torch.full((V,), float('-inf'))
        mask[valid] = 0.0
        table = table + mask

    def step(last, state):
        logits = table[last] + state['bias']
        return torch.log_softmax(logits, 1), {'bias': state['bias'] * 0.9, 'count': state['count'] + 1}
    return step


def run_both(V, B, beam, per_node, max_steps, seed, end_bias=0.0, valid=None):
    from models.allennlp_beamsearch import BeamSearch
    g = torch.Generator().manual_seed(seed + 1)
    start = torch.full((B,), 1, dtype=torch.int64)
    mk = lambda: {'bias': torch.randn(B, V, generator=torch.Generator().manual_seed(seed + 2)) * 0.3, 'count': torch.zeros(B, 1)}
    ref = O.beam_search(markov(V, seed, end_bias, valid), start, mk(), END, max_steps, beam, per_node)
    got = BeamSearch(END, max_steps=max_steps, beam_size=beam, per_node_beam_size=per_node).search(
        start, mk(), markov(V, seed, end_bias, valid))
    return ref, got


def check(ref, got):
    assert got[0].shape == ref[0].shape and got[0].dtype == torch.int64
    assert torch.equal(got[0], ref[0])
    assert torch.allclose(got[1], ref[1], atol=1e-5, equal_nan=True)


@pytest.mark.parametrize('V,B,beam,per_node,T', [(11, 3, 3, 2, 6), (37, 2, 5, 5, 7), (9, 4, 1, 1, 5), (6, 1, 4, 6, 4)])
def test_matches_reference_algorithm(V, B, beam, per_node, T):
    check(*run_both(V, B, beam, per_node, T, seed=V + beam))


def test_early_stop_returns_the_steps_actually_taken():
    ref, got = run_both(V=8, B=3, beam=3, per_node=3, max_steps=12, seed=5, end_bias=25.0)
    assert ref[0].shape[2] < 12                                   # every beam hit <end>: the loop broke early
    check(ref, got)


def test_beam_one_all_end_early_return_warns():
    from models.allennlp_beamsearch import BeamSearch
    step = markov(7, 3, end_bias=60.0)
    st0 = {'bias': torch.zeros(2, 7), 'count': torch.zeros(2, 1)}
    with pytest.warns(RuntimeWarning, match='Empty sequences'):
        pred, lp = BeamSearch(END, max_steps=5, beam_size=1).search(torch.ones(2, dtype=torch.int64), st0, step)
    assert pred.shape == (2, 1, 1) and bool((pred == END).all()) and lp.shape == (2, 1)


def test_per_node_beam_larger_than_vocab_is_a_configuration_error():
    from models.allennlp_beamsearch import BeamSearch
    st0 = {'bias': torch.zeros(2, 4), 'count': torch.zeros(2, 1)}
    with pytest.raises(ConfigurationError, match='too small relative to per_node_beam_size'):
        BeamSearch(END, max_steps=4, beam_size=3, per_node_beam_size=5).search(torch.ones(2, dtype=torch.int64), st0, markov(4, 1))


def test_fewer_valid_transitions_than_beams_warns_about_infinite_scores():
    from models.allennlp_beamsearch import BeamSearch
    V = 9
    step = markov(V, 11, valid=[3])                                 # a single reachable token: one finite path, four beams
    st0 = {'bias': torch.zeros(1, V), 'count': torch.zeros(1, 1)}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        pred, lp = BeamSearch(END, max_steps=3, beam_size=4).search(torch.ones(1, dtype=torch.int64), st0, step)
    assert any('Infinite log probabilities' in str(x.message) for x in w)
    assert torch.isinf(lp).any() and pred.shape == (1, 4, 3)


@settings(max_examples=25, deadline=None)
@given(V=st.integers(5, 24), B=st.integers(1, 4), beam=st.integers(1, 5), per_node=st.integers(1, 5), T=st.integers(2, 8),
       seed=st.integers(0, 10 ** 6), end_bias=st.sampled_from([0.0, 3.0, 12.0]))
def test_random_shapes_match_reference_algorithm(V, B, beam, per_node, T, seed, end_bias):
    ops.set_backend(CpuEmulBackend())
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref, got = run_both(V, B, beam, per_node, T, seed, end_bias)
    check(ref, got)
