"""CPU oracle for the iterative-inference hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain PyTorch-CPU / numpy
restatement of the reference algorithm (adri-romsor/iterative_inference_segm,
Theano/Lasagne) and exists to check the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package
(``iterative_inference_segm_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and
Theano/Lasagne (pinned at Theano ddafc3e2, Lasagne 45bb5689, reference
README.md:142) cannot be installed here, so the oracle could not be checked
against outputs of the reference itself.  What pins it instead: the shape
tables derivable from the reference code, hand-computed metric examples, the
Lasagne semantics restated in SURVEY.md App. A, and an fp64 re-run of the same
restatement (tests/test_oracle.py).
"""
from . import lasagne_semantics, nets, metrics, loop, weights  # noqa: F401
