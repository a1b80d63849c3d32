"""Host-side checks (no GPU) of the rows SURVEY.md section 8f adds around the hot path: spectral-normalised coloring layers
(f3), checkpoint name interchange (f4), the block / generator / discriminator assembly (f1) and the objectives (f2)."""
import numpy as np
import pytest
import torch


def test_max_singular_val_converges_to_the_spectral_norm():
    """gan/unit_tests.py:136-160 criterion: after enough power iterations sigma is the largest singular value."""
    import wc_gan_b200 as W
    g = torch.Generator().manual_seed(0)
    w = torch.randn(3, 12, 7, generator=g, dtype=torch.float64)
    u = torch.randn(3, 12, generator=g, dtype=torch.float64)
    sigma, u_bar = W.max_singular_val(w, u, ip=200)
    ref = torch.linalg.svdvals(w)[:, 0]
    assert torch.allclose(sigma, ref, rtol=1e-8)
    assert torch.allclose(u_bar.norm(dim=-1), torch.ones(3, dtype=torch.float64))
    # one iteration from the converged u reproduces sigma (the stateful update of the layers)
    sigma2, _ = W.max_singular_val(w, u_bar, ip=1)
    assert torch.allclose(sigma2, ref, rtol=1e-8)


@pytest.mark.parametrize('renormalize', [False, True])
def test_sn_coloring_layers_normalise_per_matrix(renormalize):
    import wc_gan_b200 as W
    torch.manual_seed(0)
    layer = W.SNConditionalConv11(filters=6, number_of_classes=4, in_channels=6, spectral_iterations=50, renormalize=renormalize)
    w = layer.weights()['Wc']
    assert tuple(w.shape) == (4, 6, 6)
    if renormalize:
        assert abs(float(torch.linalg.svdvals(w.reshape(-1, 6))[0]) - 1.0) < 1e-4
    else:
        assert torch.allclose(torch.linalg.svdvals(w)[:, 0], torch.ones(4), atol=1e-4)
    assert 'largest_singular_value' in dict(layer.named_buffers())
    f = W.SNFactorizedConv11(filters=6, number_of_classes=4, filters_emb=3, in_channels=6, spectral_iterations=50)
    assert torch.allclose(torch.linalg.svdvals(f.weights()['D'])[:, 0], torch.ones(3), atol=1e-4)
    u = W.SNConv11(filters=6, in_channels=6, spectral_iterations=50)
    assert abs(float(torch.linalg.svdvals(u.weights()['Wu'])[0]) - 1.0) < 1e-4
    # gradients flow to the raw kernel through the division by sigma (fully_diff_spectral=True)
    layer.weights()['Wc'].sum().backward()
    assert layer.kernel.grad is not None and torch.isfinite(layer.kernel.grad).all()


def test_generator_assembly_and_keras_names():
    import wc_gan_b200 as W
    G = W.make_generator(block_sizes=(16, 16), resamples=('UP', 'UP'), first_block_shape=(4, 4, 16), number_of_classes=5,
                         block_norm='d', block_after_norm='ufconv', filters_emb=3, last_norm='d', last_after_norm='uconv')
    stacks = G.wc_layers()
    assert len(stacks) == 5                                  # 2 blocks x (bn1, bn2) + the final norm  (generator.py:141-154)
    assert all(s.activation == 'relu' for s in stacks)       # the ReLU after every WC layer is fused
    assert len(list(G.parameters())) > 0 and all(s.built for s in stacks)
    names = W.checkpoint.keras_names(G)
    assert names['blocks.0.bn1.npart.moving_variance'] == 'Generator.0_bn1_npart/moving_variance:0'
    assert names['blocks.1.bn2.repart_c.class_matrix'] == 'Generator.1_bn2_repart_c/class_matrix:0'
    assert names['last_norm.repart.kernel'] == 'Generator.BN.Final_repart/kernel:0'
    # round trip through the Keras-named mapping
    w = W.checkpoint.export_keras_weights(G)
    assert w['Generator.0_bn1_npart/moving_mean:0'].shape == (16, 1) and w['Generator.0_bn1_repart_c/kernel:0'].shape == (3, 1, 1, 16, 16)
    G2 = W.make_generator(block_sizes=(16, 16), resamples=('UP', 'UP'), first_block_shape=(4, 4, 16), number_of_classes=5,
                          block_norm='d', block_after_norm='ufconv', filters_emb=3, last_norm='d', last_after_norm='uconv')
    W.checkpoint.load_keras_weights(G2, w)
    for k in names:
        assert torch.equal(G.state_dict()[k], G2.state_dict()[k])
    with pytest.raises(KeyError):
        W.checkpoint.load_keras_weights(G2, {})


def test_discriminator_and_objectives_on_cpu(tmp_path):
    """The discriminator has no WC layer by default (norm='n', discriminator.py:19): it runs on CPU."""
    import wc_gan_b200 as W
    D = W.make_discriminator((16, 16, 3), block_sizes=(8, 8, 8), resamples=('DOWN', 'DOWN', 'SAME'), number_of_classes=5,
                             type='PROJECTIVE', spectral=True, sum_pool=True)
    x = torch.randn(4, 16, 16, 3)
    cls = torch.randint(0, 5, (4, 1), dtype=torch.int32)
    out = D(x, cls)
    assert tuple(out.shape) == (4, 1)
    real, fake = torch.tensor([[2.0], [0.5]]), torch.tensor([[-2.0], [0.0]])
    assert float(W.discriminator_loss(real, fake, 'hinge')) == pytest.approx(0.25 + 0.5)    # gan/gan.py:111-115
    assert float(W.generator_loss(fake, 'hinge')) == pytest.approx(1.0)                     # :85-86
    # resblock geometry (gan/layer_utils.py:134-188): UP doubles, DOWN halves, channel change goes through the 1x1 shortcut
    up = W.ResBlock(8, (3, 3), 'UP', 12, 'b')
    assert tuple(up(torch.randn(2, 4, 4, 8)).shape) == (2, 8, 8, 12)
    down = W.ResBlock(8, (3, 3), 'DOWN', 8, 'b', is_first=True)
    assert tuple(down(torch.randn(2, 4, 4, 8)).shape) == (2, 2, 2, 8)
    dc = W.DCBlock(8, (4, 4), 'UP', 6, 'd')
    assert tuple(dc(torch.randn(2, 4, 4, 8)).shape) == (2, 8, 8, 6)
