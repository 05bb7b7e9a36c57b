import theano.tensor as T


def rectify(x):
    return T.nnet.relu(x)


def linear(x):
    return x


identity = linear


def softmax(x):
    return T.nnet.softmax(x)


def sigmoid(x):
    return T.nnet.sigmoid(x)
