"""lasagne.updates stand-in (test infrastructure): the optimisers train_dae.py can select."""
from collections import OrderedDict

import numpy as np
import theano
import theano.tensor as T


def get_or_compute_grads(loss_or_grads, params):
    if isinstance(loss_or_grads, list):
        assert len(loss_or_grads) == len(params)
        return loss_or_grads
    return theano.grad(loss_or_grads, params)


def sgd(loss_or_grads, params, learning_rate):
    grads = get_or_compute_grads(loss_or_grads, params)
    return OrderedDict((param, param - learning_rate * grad) for param, grad in zip(params, grads))


def rmsprop(loss_or_grads, params, learning_rate=1.0, rho=0.9, epsilon=1e-6):
    grads = get_or_compute_grads(loss_or_grads, params)
    updates = OrderedDict()
    for param, grad in zip(params, grads):
        value = param.get_value(borrow=True)
        accu = theano.shared(np.zeros(value.shape, dtype=value.dtype))
        accu_new = rho * accu + (1 - rho) * grad ** 2
        updates[accu] = accu_new
        updates[param] = param - (learning_rate * grad / T.sqrt(accu_new + epsilon))
    return updates


def adam(loss_or_grads, params, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
    all_grads = get_or_compute_grads(loss_or_grads, params)
    t_prev = theano.shared(np.float32(0.))
    updates = OrderedDict()
    t = t_prev + 1
    a_t = learning_rate * T.sqrt(1 - float(beta2) ** t) / (1 - float(beta1) ** t)
    for param, g_t in zip(params, all_grads):
        value = param.get_value(borrow=True)
        m_prev = theano.shared(np.zeros(value.shape, dtype=value.dtype))
        v_prev = theano.shared(np.zeros(value.shape, dtype=value.dtype))
        m_t = beta1 * m_prev + (1 - beta1) * g_t
        v_t = beta2 * v_prev + (1 - beta2) * g_t ** 2
        step = a_t * m_t / (T.sqrt(v_t) + epsilon)
        updates[m_prev] = m_t
        updates[v_prev] = v_t
        updates[param] = param - step
    updates[t_prev] = t
    return updates
