"""Pin the CPU oracle (oracle/likelihood_oracle.c) to the reference: every golden state dumped
from the reference's own AVX build must be reproduced -- lnL and per-partition lnL of the first
evaluation and of the optimised `-f e` end state, root scaler counts, the first makenewz result."""
import pytest

from replay import CASES, PARTIAL_CASES, replay_case, replay_partial


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_reproduces_reference(oracle_factory, tag):
    replay_case(oracle_factory, tag)


@pytest.mark.parametrize("tag", PARTIAL_CASES)
def test_oracle_reproduces_reference_evaluate_partial(oracle_factory, tag):
    assert replay_partial(oracle_factory, tag) >= 160


def test_golden_headline_values():
    """The frozen goldens are the values SURVEY.md 6 / BASELINE.md 2 quote for the reference."""
    from golden_util import load_json
    assert load_json("dna49.first.json")["likelihood"] == -19695.168878012522
    assert load_json("dna49.final.json")["likelihood"] == -16235.408714647221
    assert load_json("dna49p.final.json")["likelihood"] == -16205.671990110859
    assert load_json("aa140.first.json")["likelihood"] == -129315.0588145853
    assert load_json("aa140.final.json")["likelihood"] == -120967.43578720633
    assert abs(load_json("dna49pM.final.json")["likelihood"] - -16035.202133) < 1e-6
