"""Runs the REFERENCE's own Python sources (read where they lie under /root/reference, never copied) in this container.

TEST INFRASTRUCTURE, like the rest of oracle/: only tests/ and tests/golden/make_reference_golden.py use it.

The reference is Python-2 Theano / Lasagne code (README.md:17-22) and neither library is installable here.  What its
sources need from them is small: a symbolic-expression front end, `theano.function`, `T.grad`, and about twenty Lasagne layer
classes.  `stubs/theano` and `stubs/lasagne` provide exactly that surface on torch CPU float32 tensors (a lazy expression
graph evaluated per `theano.function` call; `T.grad` is torch autograd; the max-pool gradient follows Theano's CPU
MaxPoolGrad, which credits every element equal to the window maximum).  `py2import` loads the reference's modules from
/root/reference through an import hook that rewrites Python-2 print statements and `/` on integers in memory.

With these, tests/golden/make_reference_golden.py executes the reference's model builders (models/fcn8.py, models/DAE_h.py,
models/fcn_down.py, models/fcn_up.py, models/model_helpers.py, models/contextmod_dae.py, models/fcn8_dae.py,
layers/mylayers.py), its metrics (metrics.py) and its drivers (iterative_inference.py:inference,
iterative_inference_valid.py:inference) UNMODIFIED and stores their outputs as fixtures under tests/golden/ref_*.npz.
The oracle restatement (oracle/nets.py, loop.py, metrics.py) is then checked against those fixtures (tests/test_oracle.py),
which pins layer wiring, parameter order, concatenation order, crop offsets, the DePool2D mask rule, the loop body and
the metrics to the reference's code.  What stays restated: the Lasagne layer arithmetic inside stubs/lasagne (written
independently of oracle/lasagne_semantics.py, e.g. the dilated and transposed convolutions go through
torch.nn.grad.conv2d_weight / conv2d_input, the way Lasagne defines them).
"""
