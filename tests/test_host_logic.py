"""Host-side control flow of the autograd.Function with the kernel launches stubbed out: which C-ABI calls are
enqueued, with which WcLayerCall contents, for each create_norm variant.  Runs on CPU so that Python-level mistakes are caught
before a GPU run; no arithmetic happens here (outputs are uninitialised)."""
import pytest
import torch


@pytest.fixture
def stubbed(monkeypatch):
    from wc_gan_b200 import functional as F
    calls = []
    monkeypatch.setattr(F, '_call', lambda name, *a: calls.append((name, a)))
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
    if norm == 'n' and after_norm == 'n':                         # generator.py:19-20,79-80: the identity, no pass at all
        assert stubbed == [] and y is x
        return
    assert [n for n, _ in stubbed] == ['wc_layer_fwd']            # one foreign call per direction
    call = stubbed[0][1][0]._obj
    assert call.phases == 3 and call.training == 1
    assert call.whiten == (norm != 'n') and call.renorm == (norm == 'dr')
    assert call.K == (3 if after_norm in ('cconv', 'ucconv', 'fconv', 'ufconv') else 1)
    assert bool(call.Wu) == (after_norm in ('uconv', 'ucconv', 'ufconv'))
    assert bool(call.Wc) == (after_norm in ('cconv', 'ucconv')) and bool(call.D) == (after_norm in ('fconv', 'ufconv'))
    assert call.saved_bytes >= 4 * c * c and bool(call.y) and bool(call.x)
    del stubbed[:]
    y.backward(torch.zeros_like(y))
    assert [n for n, _ in stubbed] == ['wc_layer_bwd']
    call = stubbed[0][1][0]._obj
    assert call.phases == 3 and bool(call.dx) and bool(call.dy)
    assert bool(call.red) == (norm != 'n')                       # the whitening adjoint only exists in the training phase
    assert bool(call.dWu) == (after_norm in ('uconv', 'ucconv', 'ufconv'))
    # inference phase: no statistics, no factorisation adjoint
    del stubbed[:]
    stack.eval()
    y = stack(x, cls)
    y.backward(torch.zeros_like(y))
    assert [n for n, _ in stubbed] == ['wc_layer_fwd', 'wc_layer_bwd']
    assert stubbed[0][1][0]._obj.training == 0 and not stubbed[1][1][0]._obj.red


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


def test_operand_validation(stubbed):
    """whitening_coloring() hands raw pointers to the C ABI: every operand's dtype, layout and shape is checked first
    (a mismatch used to read out of bounds silently)."""
    from wc_gan_b200.functional import WCSpec, whitening_coloring, MAX_GROUP
    x = torch.zeros(4, 2, 2, 8)
    mm, mc = torch.zeros(8, 1), torch.eye(8)
    spec = WCSpec(number_of_classes=3)
    ok = dict(spec=spec, training=True, moving_mean=mm, moving_cov=mc, cls=torch.zeros(4, 1, dtype=torch.int32))
    whitening_coloring(x, Wu=torch.zeros(8, 8), Wc=torch.zeros(3, 8, 8), bc=torch.zeros(3, 8), **ok)
    with pytest.raises(ValueError):
        whitening_coloring(x, Wu=torch.zeros(8, 8), Wc=torch.zeros(2, 8, 8), **ok)          # K mismatch
    with pytest.raises(ValueError):
        whitening_coloring(x, Wu=torch.zeros(8, 6), Wc=torch.zeros(3, 8, 8), **ok)          # Cout disagrees
    with pytest.raises(ValueError):
        whitening_coloring(x, Wc=torch.zeros(3, 8, 8), bc=torch.zeros(3, 4), **ok)          # bias shape
    with pytest.raises(ValueError):
        whitening_coloring(x, class_matrix=torch.zeros(3, 2), **ok)                          # dictionary missing
    with pytest.raises(ValueError):
        whitening_coloring(x, class_matrix=torch.zeros(2, 2), D=torch.zeros(2, 8, 8), **ok)  # class_matrix rows != K
    with pytest.raises(TypeError):
        whitening_coloring(x, Wu=torch.zeros(8, 8, dtype=torch.float64), **ok)
    with pytest.raises(ValueError):
        whitening_coloring(x, Wu=torch.zeros(8, 8), spec=spec, training=True, moving_mean=mm, moving_cov=torch.eye(4),
                           cls=ok['cls'])                                                    # moving_cov size
    with pytest.raises(ValueError):
        whitening_coloring(x, Wu=torch.zeros(8, 8), spec=spec, training=True, moving_mean=mm,
                           moving_cov=torch.eye(16)[::2, ::2], cls=ok['cls'])                # in-place buffer not contiguous
    big = torch.zeros(1, 1, 2, 2 * MAX_GROUP)
    with pytest.raises(ValueError, match='group_size'):
        whitening_coloring(big, spec=WCSpec(), training=True, moving_mean=torch.zeros(2 * MAX_GROUP),
                           moving_cov=torch.eye(2 * MAX_GROUP))                              # g = C > 256: named limit


def test_peer_exchange_is_only_offered_on_cuda(monkeypatch):
    """The peer-memory transport is never chosen for CPU tensors or when WC_B200_SYNC=nccl asks for host collectives;
    neither case touches torch.distributed."""
    import torch
    from wc_gan_b200 import sync
    assert sync.peer_exchange(None, 8, 8, torch.device('cpu')) is None
    monkeypatch.setenv('WC_B200_SYNC', 'nccl')
    assert sync.peer_exchange(None, 256, 256, torch.device('cuda', 0)) is None
    assert sync.record_numel(8, 4) == 8 + 2 * 16 + 1
