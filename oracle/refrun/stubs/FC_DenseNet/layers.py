"""The four helpers models/FCDenseNet.py:12 imports, restated from the published FC-DenseNet repository / arXiv 1611.09326 on
the Lasagne stand-in (test infrastructure).  The package is NOT in /root/reference, so these four functions are a
restatement; the network that calls them -- models/FCDenseNet.py:Network and build_fcdensenet -- is the reference's own code."""
from lasagne.init import HeUniform
from lasagne.layers import (BatchNormLayer, ConcatLayer, Conv2DLayer, Deconv2DLayer, DimshuffleLayer, DropoutLayer, NonlinearityLayer,
                            Pool2DLayer, ReshapeLayer)
from lasagne.nonlinearities import linear, softmax


def BN_ReLU_Conv(inputs, n_filters, filter_size=3, dropout_p=0.2):
    l = NonlinearityLayer(BatchNormLayer(inputs))
    l = Conv2DLayer(l, n_filters, filter_size, pad='same', W=HeUniform(gain='relu'), nonlinearity=linear, flip_filters=False)
    if dropout_p != 0.0:
        l = DropoutLayer(l, dropout_p)
    return l


def TransitionDown(inputs, n_filters, dropout_p=0.2):
    l = BN_ReLU_Conv(inputs, n_filters, filter_size=1, dropout_p=dropout_p)
    l = Pool2DLayer(l, 2, mode='max')
    return l


def TransitionUp(skip_connection, block_to_upsample, n_filters_keep):
    l = ConcatLayer(block_to_upsample)
    l = Deconv2DLayer(l, n_filters_keep, filter_size=3, stride=2, crop='valid', W=HeUniform(gain='relu'), nonlinearity=linear)
    l = ConcatLayer([l, skip_connection], cropping=[None, None, 'center', 'center'])
    return l


def SoftmaxLayer(inputs, n_classes):
    l = Conv2DLayer(inputs, n_classes, filter_size=1, nonlinearity=linear, W=HeUniform(gain='relu'), pad='same', flip_filters=False, stride=1)
    l = DimshuffleLayer(l, (0, 2, 3, 1))
    l = ReshapeLayer(l, (-1, n_classes))
    l = NonlinearityLayer(l, softmax)
    return l
