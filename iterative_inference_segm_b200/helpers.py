"""helpers.py of the reference: the parts the iterative-inference scripts use."""
import numpy as np


def build_experiment_name(segm_net='fcn8', kind='fcn8', concat_h=[], optimizer='rmsprop',
                          training_loss=['crossentropy'], learning_rate=0.0001, lr_anneal=0.99,
                          data_aug=False, weight_decay=0.0001, dropout=0.5, noise=0.0, from_gt=False,
                          temperature=1.0, n_filters=64, conv_before_pool=1, skip=True,
                          additional_pool=0, unpool_type='standard', ae_h=False, path_weights='',
                          layer='probs_dimshuffle', exp_name='', bn=0):
    """Directory name of an experiment; same fields in the same order as helpers.py:118-169 so
    checkpoints saved by the reference's train_dae.py are found under the same path."""
    name = exp_name + segm_net + '_' + kind + '_' + '_'.join(concat_h)
    if kind == 'standard':
        name += '_f%sc%sp%s%s_%s' % (n_filters, conv_before_pool, additional_pool, '_skip' if skip else '', unpool_type)
    if dropout > 0.:
        name += '_dropout' + str(dropout)
    name += '_' + '_'.join(training_loss)
    name += ('_fromgt' if from_gt else '_fromfcn8') + '_z' + str(noise)
    if bool(data_aug):
        name += '_data_aug'
    if not from_gt:
        name += '_T' + str(temperature)
    name += '_%s_lr%s_anneal%s_decay%s' % (optimizer, learning_rate, lr_anneal, weight_decay)
    if len(path_weights) > 0:
        name += '_pretrained'
    if ae_h:
        name += '_PlugPlay'
    name += '_' + layer
    if bn:
        name += '_bn'
    return name


def results_values(rec, acc, jacc, nbatches):
    """(loss, acc, mean Jaccard) exactly as helpers.py:172-177 reports them."""
    with np.errstate(divide='ignore', invalid='ignore'):
        jacc_mean = np.nanmean(jacc[0, :] / jacc[1, :])
    return rec / nbatches, acc / nbatches, jacc_mean


def print_results(st, rec, acc, jacc, nbatches):
    loss, a, jm = results_values(rec, acc, jacc, nbatches)
    print(st)
    print('    Loss: ' + str(loss))
    print('    Acc: ' + str(a))
    print('    Jaccard: ' + str(jm))
