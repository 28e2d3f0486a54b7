"""GPU parity of the optimizer side of the training step (BASELINE config 4, train.py:109-110) against
torch.optim.Adamax and nn.utils.clip_grad_norm_ on the same parameters and gradients."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


SHAPES = [(), (1,), (7,), (4097,), (300, 17), (1024, 1024), (3,), (64, 4096), (5, 5, 5)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.rand(s, generator=g) - 0.5).cuda()) for s in SHAPES]


def _set_grads(params, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    for p in params:
        p.grad = ((torch.rand(p.shape, generator=g) - 0.5) * scale).cuda()


def _maxrel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_adamax_and_clip_match_torch(weight_decay):
    from vqa_collection_b200 import optim
    ours, ref = _params(1), _params(1)
    groups = lambda ps: [{"params": ps[:4]}, {"params": ps[4:], "lr": 0.004}]          # train.py:54-56 param groups
    o1 = optim.Adamax(groups(ours), lr=0.002, weight_decay=weight_decay)
    o2 = torch.optim.Adamax(groups(ref), lr=0.002, weight_decay=weight_decay)
    for it in range(6):
        big = 30.0 if it % 2 == 0 else 1e-3                           # clipped and not clipped steps
        _set_grads(ours, 100 + it, big)
        _set_grads(ref, 100 + it, big)
        n1 = optim.clip_grad_norm_(ours, 0.25)
        n2 = torch.nn.utils.clip_grad_norm_(ref, 0.25)
        assert n1.is_cuda and n1.dim() == 0 and abs(float(n1) - float(n2)) <= 2e-6 * float(n2)
        for a, b in zip(ours, ref):
            assert _maxrel(a.grad, b.grad) < 2e-6
        o1.step()
        o2.step()
        o1.zero_grad()
        o2.zero_grad()
        for a, b in zip(ours, ref):
            assert _maxrel(a.detach(), b.detach()) < 2e-6, (it, tuple(a.shape))
    for a, b in zip(ours, ref):
        s1, s2 = o1.state[a], o2.state[b]
        assert float(s1["step"]) == float(s2["step"]) == 6
        assert _maxrel(s1["exp_avg"], s2["exp_avg"]) < 2e-6 and _maxrel(s1["exp_inf"], s2["exp_inf"]) < 2e-6


def test_adamax_state_dict_round_trips_with_torch_and_steplr():
    from vqa_collection_b200 import optim
    ours, ref = _params(2), _params(2)
    o1, o2 = optim.Adamax(ours, lr=0.002), torch.optim.Adamax(ref, lr=0.002)
    sched = torch.optim.lr_scheduler.StepLR(o1, step_size=1, gamma=0.5)                 # train.py:58
    for it in range(2):
        _set_grads(ours, 7 + it)
        _set_grads(ref, 7 + it)
        o1.step()
        o2.step()
    # hand the fused optimizer's state to torch's and continue there: same trajectory as torch alone
    o3 = torch.optim.Adamax(ours, lr=0.002)
    o3.load_state_dict(copy.deepcopy(o1.state_dict()))
    _set_grads(ours, 50)
    _set_grads(ref, 50)
    o3.step()
    o2.step()
    for a, b in zip(ours, ref):
        assert _maxrel(a.detach(), b.detach()) < 2e-6
    # and back
    o4 = optim.Adamax(ref, lr=0.002)
    o4.load_state_dict(copy.deepcopy(o2.state_dict()))
    _set_grads(ours, 51)
    _set_grads(ref, 51)
    o3.step()
    o4.step()
    for a, b in zip(ours, ref):
        assert _maxrel(a.detach(), b.detach()) < 2e-6
    sched.step()
    assert abs(o1.param_groups[0]["lr"] - 0.001) < 1e-12


def test_adamax_fused_grad_scale_and_skipped_params():
    from vqa_collection_b200 import optim
    ours, ref = _params(3), _params(3)
    o1, o2 = optim.Adamax(ours, lr=0.01), torch.optim.Adamax(ref, lr=0.01)
    _set_grads(ours, 9)
    _set_grads(ref, 9)
    ours[2].grad = None                                                   # a parameter without a gradient is left alone
    ref[2].grad = None
    before = ours[2].detach().clone()
    for p in ref:
        if p.grad is not None:
            p.grad.mul_(0.125)
    o1.step(grad_scale=torch.tensor([0.125], device="cuda"))
    o2.step()
    for a, b in zip(ours, ref):
        assert _maxrel(a.detach(), b.detach()) < 2e-6
    assert torch.equal(ours[2].detach(), before) and len(o1.state[ours[2]]) == 0
