"""CPU oracle for the iterative-inference hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain PyTorch-CPU / numpy
restatement of the reference algorithm (adri-romsor/iterative_inference_segm,
Theano/Lasagne) and exists to check the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package
(``iterative_inference_segm_b200``) never imports it and has no CPU fallback.

PARITY PINNED BY EXECUTING THE REFERENCE (round 2).  The reference ships no tests, golden vectors or fixtures, and
Theano / Lasagne (Theano ddafc3e2, Lasagne 45bb5689, reference README.md:142) cannot be installed here.  `oracle/refrun`
therefore provides stand-ins for the small part of both libraries the reference uses, plus a Python-2 import hook, and
`tests/golden/make_reference_golden.py` runs the reference's OWN, unmodified sources with them: its drivers
(iterative_inference.py:inference, iterative_inference_valid.py:inference), model builders (models/fcn8.py, DAE_h.py,
fcn_down.py, fcn_up.py, model_helpers.py, contextmod_dae.py, fcn8_dae.py, layers/mylayers.py), metrics.py and, for the
training step, train_dae.py:train.  Their outputs are committed as tests/golden/ref_*.npz and this restatement replays
every one of them to 3e-7 (tests/test_oracle.py::test_oracle_vs_reference_run; integer matrices exactly).  What remains
restated rather than executed: the arithmetic of the Lasagne layers inside oracle/refrun/stubs/lasagne (written
independently of oracle/lasagne_semantics.py), Theano's CPU MaxPoolGrad tie rule, and the four layer helpers FC-DenseNet103
imports from the absent FC_DenseNet package (the network that calls them, models/FCDenseNet.py, is executed).  Also pinned by hand: the shape tables, hand-computed metric examples and
an fp64 re-run (tests/test_oracle.py).
"""
from . import lasagne_semantics, nets, metrics, loop, weights  # noqa: F401
