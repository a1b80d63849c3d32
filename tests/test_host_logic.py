"""Host-side control flow of the autograd.Function with the kernel launches stubbed out: which C-ABI stages are
enqueued, in which order, for each create_norm variant.  Runs on CPU so that Python-level mistakes are caught
before a GPU run; no arithmetic happens here (outputs are uninitialised)."""
import pytest
import torch


@pytest.fixture
def stubbed(monkeypatch):
    from wc_gan_b200 import functional as F
    calls = []
    monkeypatch.setattr(F, '_call', lambda name, *a: calls.append(name))
    monkeypatch.setattr(F, '_need_cuda', lambda t, name: None)
    monkeypatch.setattr(F, '_stream', lambda: 0)
    return calls


@pytest.mark.parametrize('norm', ['d', 'dr', 'n'])
@pytest.mark.parametrize('after_norm', ['uconv', 'cconv', 'ucconv', 'fconv', 'ufconv', 'n'])
@pytest.mark.parametrize('c', [24, 128])
def test_stage_sequence(stubbed, norm, after_norm, c):
    import wc_gan_b200 as W
    stack = W.create_norm(norm, after_norm, number_of_classes=3, filters_emb=2)(axis=-1, name='G.0_bn1')
    x = torch.zeros(4, 2, 2, c, requires_grad=True)
    cls = torch.zeros(4, 1, dtype=torch.int32)
    stack.train()
    y = stack(x, cls)
    assert y.shape == x.shape
    fwd = list(stubbed)
    tc = c == 128
    expect = (['wc_stats_fwd', 'wc_factor_fwd'] if norm != 'n' else []) + ['wc_compose_fwd'] + (['wc_pack_fwd'] if tc else []) + ['wc_apply_fwd']
    assert fwd == expect
    del stubbed[:]
    y.backward(torch.zeros_like(y))
    bwd = list(stubbed)
    expect = ['wc_reduce_bwd', 'wc_color_bwd'] + (['wc_factor_bwd'] if norm != 'n' else []) + (['wc_pack_bwd'] if tc else []) + ['wc_apply_bwd']
    assert bwd == expect
    # inference phase: no statistics, no factorisation adjoint
    del stubbed[:]
    stack.eval()
    y = stack(x, cls)
    y.backward(torch.zeros_like(y))
    assert 'wc_stats_fwd' not in stubbed and 'wc_factor_bwd' not in stubbed
    assert ('wc_factor_fwd' in stubbed) == (norm != 'n')


def test_label_validation(stubbed):
    import wc_gan_b200 as W
    stack = W.create_norm('d', 'cconv', number_of_classes=3)(axis=-1, name='n')
    x = torch.zeros(4, 2, 2, 8)
    with pytest.raises(ValueError):
        stack(x)                                             # conditional layer without labels
    with pytest.raises(ValueError):
        stack(x, torch.zeros(3, 1, dtype=torch.int32))       # one label per sample
    with pytest.raises(ValueError):
        W.DecorelationNormalization(group_size=3)(x)         # channels not divisible by the group size
