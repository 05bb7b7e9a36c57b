"""Stand-in for the part of Lasagne (0.2.dev1, README.md:17-22 of the reference) that the reference's sources use.
Test infrastructure; see oracle/refrun/__init__.py.  Layer arithmetic is restated from Lasagne's published layers and
written independently of oracle/lasagne_semantics.py."""
from . import init, layers, nonlinearities, objectives, random, regularization, updates, utils  # noqa: F401
