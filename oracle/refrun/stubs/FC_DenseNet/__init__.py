"""Stand-in for the absent dependency `FC_DenseNet` (SimJeg/FC-DenseNet; models/FCDenseNet.py:12 of the reference imports four
layer helpers from it).  Test infrastructure; see oracle/refrun/__init__.py."""
