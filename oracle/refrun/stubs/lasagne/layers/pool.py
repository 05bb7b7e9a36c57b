from . import MaxPool2DLayer, Pool2DLayer, Upscale2DLayer, pool_output_length  # noqa: F401
