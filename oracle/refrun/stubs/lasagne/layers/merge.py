from . import ConcatLayer, ElemwiseMergeLayer, ElemwiseSumLayer, autocrop, autocrop_array_shapes  # noqa: F401
