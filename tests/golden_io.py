"""Read the committed fixtures of ``tests/golden`` back into (op, tensors, params, expected)."""
import glob
import json
import os
from collections import OrderedDict

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
# order in which each op's adapter expects its tensors (None entries are restored from meta)
INPUT_ORDER = {
    'gae': ['value', 'next_value', 'reward', 'done', 'traj_flag'],
    'ppo': ['logit_new', 'logit_old', 'action', 'value_new', 'value_old', 'adv', 'return_', 'weight',
            'logit_pretrained'],
    'ppo_policy': ['logit_new', 'logit_old', 'action', 'adv', 'weight', 'logit_pretrained'],
    'ppo_value': ['value_new', 'value_old', 'return_', 'weight'],
    'ppoc': ['mu_new', 'sigma_new', 'mu_old', 'sigma_old', 'action', 'value_new', 'value_old', 'adv', 'return_', 'weight',
             'mu_pretrained', 'sigma_pretrained'],
    'a2c': ['logit', 'action', 'value', 'adv', 'return_', 'weight'],
    'vtc': ['mu_target', 'sigma_target', 'mu_behaviour', 'sigma_behaviour', 'action', 'value', 'reward', 'weight'],
    'qntd': ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'weight', 'value_gamma'],
    'qntd_rescale': ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'weight', 'value_gamma'],
    'q1td': ['q', 'next_q', 'act', 'next_act', 'reward', 'done', 'weight'],
    'v1td': ['v', 'next_v', 'reward', 'done', 'weight'],
    'vntd': ['v', 'next_n_v', 'reward', 'done', 'weight', 'value_gamma'],
    'bdq': ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'weight', 'value_gamma'],
    'qseq': ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'weight', 'value_gamma'],
    'd1td': ['dist', 'next_dist', 'act', 'next_act', 'reward', 'done', 'weight'],
    'dntd': ['dist', 'next_n_dist', 'act', 'next_n_act', 'reward', 'done', 'weight', 'value_gamma'],
    'td_lambda': ['value', 'reward', 'weight'],
    'upgo': ['target_output', 'action', 'rhos', 'rewards', 'bootstrap_values', 'mask'],
    'vtrace': ['target_output', 'behaviour_output', 'action', 'value', 'reward', 'weight'],
    'qrdqn': ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'tau', 'weight', 'value_gamma'],
    'iqn': ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'replay_quantiles', 'weight', 'value_gamma'],
    'happo': ['logit_new', 'logit_old', 'action', 'value_new', 'value_old', 'adv', 'return_', 'weight', 'factor'],
    'acer': ['q_values', 'q_retraces', 'v_pred', 'target_logit', 'actions', 'ratio', 'avg_logit', 'actor_gradient'],
    'happoc': ['mu_new', 'sigma_new', 'mu_old', 'sigma_old', 'action', 'value_new', 'value_old', 'adv', 'return_', 'weight',
               'factor'],
    'ppg': ['logit_new', 'logit_old', 'action', 'value_new', 'value_old', 'return_', 'weight'],
    'retrace': ['q_values', 'v_pred', 'rewards', 'actions', 'weights', 'ratio'],
    'fqf': ['q', 'next_n_q', 'action', 'next_n_action', 'reward', 'done', 'quantiles_hats', 'weight', 'value_gamma'],
}


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    op = meta['op']
    tensors = OrderedDict()
    for k in INPUT_ORDER[op]:
        if 'in_' + k in z.files:
            tensors[k] = torch.from_numpy(z['in_' + k].copy())
        elif k in meta['none_inputs']:
            tensors[k] = None
    params = {}
    for k, v in meta['params'].items():
        if isinstance(v, dict) and '__tensor_list__' in v:
            v = [torch.tensor(x) for x in v['__tensor_list__']]
        params[k] = v
    expected = OrderedDict((k, z[k]) for k in z.files if k.startswith('out_') or k.startswith('grad_'))
    return op, tensors, params, expected
