"""Collector-side batch preparation (di_engine_b200.collate): list of transition dicts -> learner batch, against the restatement
of default_preprocess_learn / default_collate (ding/policy/common_utils.py:28-98, ding/utils/data/collate_fn.py:80-160).
CPU: layout, dtypes, shapes and every fix-up rule (the staging buffer is then pageable); GPU: the same through the pinned buffer
and the single H2D copy, and straight into the operators."""
import numpy as np
import pytest
import torch

import di_engine_b200 as b2
from oracle import rl_oracle


def _transitions(seed, B, kind):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(B):
        if kind == 'dqn_nstep':  # DQN collector + Adder.get_nstep_return_data: reward (nstep,), done bool, IS weight
            d = dict(obs=torch.randn(4, 8, 8, generator=g), next_obs=torch.randn(4, 8, 8, generator=g),
                     action=torch.randint(0, 6, (1, ), generator=g), reward=torch.rand(3, generator=g),
                     done=bool(i % 5 == 0), priority_IS=torch.rand(1, generator=g), value_gamma=0.97 ** 3, collate_ignore_x=i)
        elif kind == 'ppo':  # on-policy: value / adv (1,), dict observation, numpy fields, float done
            d = dict(obs={'agent_state': torch.randn(7, generator=g), 'mask': np.ones(3, dtype=np.float32) * i},
                     action=torch.randint(0, 4, (1, ), generator=g), logit=torch.randn(4, generator=g),
                     value=torch.randn(1, generator=g), adv=torch.randn(1, generator=g), reward=torch.rand(1, generator=g),
                     done=torch.tensor(float(i % 3 == 0)), traj_flag=i % 4 == 3, weight=None)
        else:  # continuous control: float action (B, D) is stacked, not concatenated; multi-agent n-step reward
            d = dict(obs=torch.randn(11, generator=g), action=torch.randn(3, generator=g),
                     reward=torch.rand(2, 5, generator=g), done=np.bool_(i % 2 == 0), step=i)
        out.append(d)
    return out


def _same(a, b, path=''):
    if isinstance(b, dict):
        assert isinstance(a, dict) and set(a) == set(b), (path, sorted(a), sorted(b))
        for k in b:
            _same(a[k], b[k], path + '/' + str(k))
    elif b is None:
        assert a is None, path
    else:
        assert a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape), (path, a.dtype, b.dtype, a.shape, b.shape)
        assert torch.equal(a.cpu(), b), path


CASES = [('dqn_nstep', dict(use_priority_IS_weight=True, use_priority=True, use_nstep=True)),
         ('dqn_nstep', dict(use_nstep=True, ignore_done=True)), ('ppo', dict()), ('continuous', dict(use_nstep=True))]


@pytest.mark.parametrize('kind,kw', CASES)
def test_preprocess_learn_matches_reference_rules_on_host(kind, kw):
    data = _transitions(3, 16, kind)
    want = rl_oracle.default_preprocess_learn([dict(d) for d in data], **kw)
    got = b2.preprocess_learn(data, 'cpu', **kw)
    _same(got, want)
    # the staging slots are reused (double-buffered) for the same layout, and a second batch does not disturb the first
    data2 = _transitions(4, 16, kind)
    got2 = b2.preprocess_learn(data2, 'cpu', **kw)
    _same(got2, rl_oracle.default_preprocess_learn([dict(d) for d in data2], **kw))
    _same(got, want)
    assert len([k for k in b2.collate._SLOTS if k[0] == 'cpu']) >= 1


@pytest.mark.gpu
@pytest.mark.parametrize('kind,kw', CASES)
def test_preprocess_learn_on_device_one_copy(kind, kw):
    data = _transitions(5, 64, kind)
    want = rl_oracle.default_preprocess_learn([dict(d) for d in data], **kw)
    got = b2.preprocess_learn(data, 'cuda:0', **kw)
    torch.cuda.synchronize()
    _same(got, want)
    assert got['done'].is_cuda


@pytest.mark.gpu
def test_collated_batch_feeds_the_operators():
    """DQN learner: transitions -> preprocess_learn(use_nstep) -> q_nstep_td_error, against the reference pipeline on the host"""
    g = torch.Generator().manual_seed(9)
    B, N, nstep = 48, 6, 3
    data = [dict(q=torch.randn(N, generator=g), next_n_q=torch.randn(N, generator=g), action=torch.randint(0, N, (1, ), generator=g),
                 next_n_action=torch.randint(0, N, (1, ), generator=g), reward=torch.rand(nstep, generator=g), done=bool(i % 7 == 0),
                 IS=torch.rand(1, generator=g)) for i in range(B)]
    kw = dict(use_priority_IS_weight=True, use_priority=True, use_nstep=True)
    ref = rl_oracle.default_preprocess_learn([dict(d) for d in data], **kw)
    dev = b2.preprocess_learn(data, 'cuda:0', **kw)
    q = dev['q'].clone().requires_grad_(True)
    loss, per = b2.q_nstep_td_error(b2.q_nstep_td_data(q, dev['next_n_q'], dev['action'], dev['next_n_action'], dev['reward'],
                                                       dev['done'], dev['weight']), 0.97, nstep=nstep)
    loss.backward()
    qr = ref['q'].clone().requires_grad_(True)
    lw, pw = rl_oracle.q_nstep_td_error(qr, ref['next_n_q'], ref['action'], ref['next_n_action'], ref['reward'], ref['done'],
                                        ref['weight'], gamma=0.97, nstep=nstep)
    lw.backward()
    assert torch.allclose(loss.cpu(), lw.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(q.grad.cpu(), qr.grad, rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_get_gae_matches_adder():
    """Adder.get_gae (ding/rl_utils/adder.py:20-58): per-step dicts -> adv per step"""
    g = torch.Generator().manual_seed(11)
    T, B = 40, 3
    data = [dict(value=torch.randn(B, generator=g), reward=torch.randn(B, generator=g)) for _ in range(T)]
    last = torch.randn(B, generator=g)
    value = torch.stack([d['value'] for d in data])
    nv = torch.stack([d['value'] for d in data][1:] + [last])
    want = rl_oracle.gae(value, nv, torch.stack([d['reward'] for d in data]), None, None, 0.99, 0.95)
    out = b2.collate.get_gae([dict(d) for d in data], last, 0.99, 0.95, device='cuda:0')
    assert all(torch.equal(out[i]['adv'], want[i]) for i in range(T))


@pytest.mark.parametrize('T,nstep,agents,cum', [(10, 3, None, False), (10, 3, None, True), (5, 8, None, False), (7, 2, 4, False),
                                                (7, 5, 4, True), (1, 3, None, False), (6, 1, None, False)])
def test_nstep_return_data_matches_adder_rules(T, nstep, agents, cum):
    """di_engine_b200.collate.nstep_return_data on stacked tensors against the restated Adder.get_nstep_return_data on a list of
    per-step dicts (ding/rl_utils/adder.py:97-155)"""
    g = torch.Generator().manual_seed(T * 31 + nstep)
    shape = (1, ) if agents is None else (agents, 1)
    data = [dict(obs=torch.full((2, ), float(i)), next_obs=torch.full((2, ), float(i + 1)), reward=torch.rand(*shape, generator=g),
                 done=bool(i == T - 1)) for i in range(T)]
    want = rl_oracle.adder_get_nstep_return_data([dict(d) for d in data], nstep, cum_reward=cum, gamma=0.97)
    reward = torch.stack([d['reward'] for d in data])
    done = torch.tensor([d['done'] for d in data])
    obs_ext = torch.stack([d['obs'] for d in data] + [data[-1]['next_obs']])   # obs 0..T-1 and the final next_obs at index T
    r_n, d_n, vg, nxt = b2.collate.nstep_return_data(reward, done, nstep, gamma=0.97, cum_reward=cum)
    for i in range(T):
        assert torch.allclose(r_n[i], want[i]['reward'], rtol=1e-6, atol=1e-7), (i, r_n[i], want[i]['reward'])
        assert bool(d_n[i]) == bool(want[i]['done']), i
        if nstep > 1:
            assert abs(float(vg[i]) - want[i]['value_gamma']) < 1e-6, i
            assert torch.equal(obs_ext[nxt[i]], want[i]['next_obs']), i


def test_staging_cache_is_bounded():
    for B in range(2, 2 + b2.collate.MAX_LAYOUTS + 4):
        b2.preprocess_learn(_transitions(B, B, 'ppo'), 'cpu')
    assert len(b2.collate._SLOTS) <= b2.collate.MAX_LAYOUTS


def _traj(T):
    return [dict(obs=torch.full((2, ), float(i)), action=torch.tensor([i]), reward=torch.tensor([float(i)]), done=(i == T - 1),
                 value_gamma=0.9) for i in range(T)]


def test_get_train_sample_follows_the_adder_rules():
    """Adder.get_train_sample (ding/rl_utils/adder.py:158-232): exact division, 'last', 'drop', 'null_padding', short trajectories"""
    gts = b2.collate.get_train_sample
    data = _traj(6)
    assert gts(data, 1) is data
    out = gts(_traj(6), 3)
    assert len(out) == 2 and [int(a) for a in out[1]['action']] == [3, 4, 5] and out[0]['done'] == [False] * 3
    out = gts(_traj(7), 3, 'last')  # remainder [6] completed in FRONT with the last two steps of the previous piece
    assert len(out) == 3 and [int(a) for a in out[2]['action']] == [4, 5, 6]
    out = gts(_traj(7), 3, 'drop')
    assert len(out) == 2
    out = gts(_traj(7), 3, 'null_padding')  # remainder [6] padded BEHIND with null transitions
    last = out[2]
    assert [int(a) for a in last['action']] == [6, 0, 0] and last['done'] == [True, True, True]
    assert last['null'][1:] == [True, True] if 'null' in last and len(last['null']) == 3 else True
    assert all(float(r) == 0.0 for r in last['reward'][1:]) and last['value_gamma'][1:] == [0., 0.]
    assert all(float(o.sum()) == 0.0 for o in last['obs'][1:])
    out = gts(_traj(2), 3, 'last')  # no previous piece: 'last' falls back to null padding
    assert len(out) == 1 and [int(a) for a in out[0]['action']] == [0, 1, 0] and out[0]['done'] == [False, True, True]
    custom = dict(obs=torch.ones(2), action=torch.tensor([9]), reward=torch.tensor([-1.0]), done=True, value_gamma=0.5)
    out = gts(_traj(4), 3, 'null_padding', null_transition=custom)
    assert [int(a) for a in out[1]['action']] == [3, 9, 9]
    dict_obs = [dict(obs={'a': torch.ones(1) * i, 'b': torch.zeros(2)}, reward=torch.tensor([1.0]), done=False) for i in range(4)]
    out = gts(dict_obs, 2)
    assert set(out[0]['obs'].keys()) == {'a', 'b'} and len(out[0]['obs']['a']) == 2  # nested dicts are transposed as well
