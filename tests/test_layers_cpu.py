"""CPU: the Python facade keeps the reference's constructor / shape / assertion contract
(mirrors /root/reference/tests/test_flows.py and tests/test_distribution_layers.py; the
numeric checks live in the gpu-marked parity tests)."""
import numpy as np
import pytest
import torch

from normalizingflownetwork_b200 import (
    FLOWS,
    AffineFlow,
    GaussianKernelsLayer,
    GaussianMixtureLayer,
    InverseNormalizingFlowLayer,
    MeanFieldLayer,
)


@pytest.mark.parametrize("flow_name", ["planar", "radial", "affine"])
def test_flow_param_sizes_and_assertions(flow_name):
    batch_size = 10
    for dim in [1, 4]:
        flow_class = FLOWS[flow_name]
        with pytest.raises(AssertionError):
            flow_class(torch.ones((batch_size, flow_class.get_param_size(dim) + 1)), dim)
        flow = flow_class(torch.ones((batch_size, flow_class.get_param_size(dim))), dim)
        reference = AffineFlow(torch.ones((batch_size, AffineFlow.get_param_size(dim))), dim)
        assert flow.forward_min_event_ndims == reference.forward_min_event_ndims == 1
    assert FLOWS["planar"].get_param_size(3) == 7
    assert FLOWS["radial"].get_param_size(3) == 5
    assert FLOWS["affine"].get_param_size(3) == 6


def test_total_param_size_nf():
    layer1 = InverseNormalizingFlowLayer(("planar", "radial", "affine"), n_dims=1, trainable_base_dist=False)
    layer2 = InverseNormalizingFlowLayer(("planar", "radial", "affine"), n_dims=3, trainable_base_dist=True)
    assert layer1.get_total_param_size() == 3 + 3 + 2
    assert layer2.get_total_param_size() == (3 + 3 + 1) + (3 + 1 + 1) + (3 + 3) + (3 + 3)
    with pytest.raises(AssertionError):
        InverseNormalizingFlowLayer(("planar", "sylvester"), n_dims=1)


def test_total_param_size_mf_and_mixture():
    assert MeanFieldLayer(n_dims=10, scale=None).get_total_param_size() == 20
    assert MeanFieldLayer(n_dims=10, scale=10.0).get_total_param_size() == 10
    assert GaussianMixtureLayer(n_dims=5, n_centers=5).get_total_param_size() == 55
    assert GaussianMixtureLayer(n_dims=1, n_centers=3).get_total_param_size() == 9


def test_mixture_dist_fn_shapes():
    dist_fn = GaussianMixtureLayer._get_distribution_fn(n_dims=1, n_centers=5)
    dist = dist_fn(torch.ones((1, 15)))
    assert dist.event_shape == [1] and dist.batch_shape == [1]
    dist = dist_fn(torch.ones((3, 15)))
    assert dist.event_shape == [1] and dist.batch_shape == [3]
    with pytest.raises(ValueError):
        dist_fn(torch.ones((10, 10)))
    dist_fn = GaussianMixtureLayer._get_distribution_fn(n_dims=3, n_centers=5)
    assert dist_fn(torch.ones((3, 35))).event_shape == [3]
    with pytest.raises(ValueError):
        dist_fn(torch.ones((10, 12)))


def test_mf_dist_fn():
    dist_fn = MeanFieldLayer._get_distribution_fn(n_dims=10, scale=None)
    dist = dist_fn(torch.ones((20,)))
    assert dist.event_shape == [10] and dist.batch_shape == []
    dist = dist_fn(torch.ones((10, 20)))
    assert dist.event_shape == [10] and dist.batch_shape == [10]
    with pytest.raises(AssertionError):
        dist_fn(torch.ones((10, 19)))
    dist_fn = MeanFieldLayer._get_distribution_fn(n_dims=10, scale=10.0)
    assert dist_fn(torch.ones((1, 10))).batch_shape == [1]
    with pytest.raises(AssertionError):
        dist_fn(torch.ones((10, 9)))


def test_nf_dist_fn_shapes_and_width_assertion():
    dist_fn = InverseNormalizingFlowLayer._get_distribution_fn(
        n_dims=1, flow_types=("radial", "planar"), trainable_base_dist=False)
    dist = dist_fn(torch.ones((1, 6)))
    assert dist.event_shape == [1] and dist.batch_shape == [1]
    assert dist.event_shape == 1  # tf.TensorShape([1]) == 1, relied on by BaseEstimator.log_pdf
    dist = dist_fn(torch.ones((3, 6)))
    assert dist.batch_shape == [3]
    with pytest.raises(AssertionError):
        dist_fn(torch.ones((10, 7)))
    dist_fn = InverseNormalizingFlowLayer._get_distribution_fn(
        n_dims=2, flow_types=("radial", "planar"), trainable_base_dist=True)
    dist = dist_fn(torch.ones((3, 13)))
    assert dist.event_shape == [2] and dist.batch_shape == [3]
    with pytest.raises(AssertionError):
        dist_fn(torch.ones((10, 12)))


def test_get_bijector_order_is_reversed():
    out = InverseNormalizingFlowLayer._get_bijector(torch.zeros((10, 8)), ("planar", "radial", "affine"), 1)
    assert len(out.bijectors) == 3
    assert out.inverse_min_event_ndims == 1
    assert type(out.bijectors[0]) == FLOWS["affine"]
    assert type(out.bijectors[1]) == FLOWS["radial"]
    assert type(out.bijectors[2]) == FLOWS["planar"]
    out = InverseNormalizingFlowLayer._get_bijector(torch.zeros((10, 9)), ("planar", "radial"), 2)
    assert len(out.bijectors) == 2
    with pytest.raises(AssertionError):
        InverseNormalizingFlowLayer._get_bijector(torch.zeros((10, 8)), ("planar", "radial"), 2)


def test_gk_layer_centers_and_scales():
    y_train = np.linspace(-1, 1, 100).reshape((100, 1))
    layer = GaussianKernelsLayer(n_centers=10, n_dims=1, trainable_scale=True, init_scales=(0.3, 0.7))
    assert layer.get_total_param_size() == 20
    dist_fn = layer._get_distribution_fn()
    assert float(layer.locs.abs().sum()) == 0.0
    layer.set_center_points(y_train)
    assert float(layer.locs.abs().sum()) != 0.0
    assert layer.locs.shape == (20, 1)
    np.testing.assert_array_equal(layer.locs[:10].numpy(), layer.locs[10:].numpy())  # tiled over the scales
    dist = dist_fn(torch.ones((3, 20)))
    assert dist.event_shape == [1] and dist.batch_shape == [3]
    assert dist.sample().shape == (3, 1)
    with pytest.raises(AssertionError):
        dist_fn(torch.ones((10, 19)))
    # bandwidth = softplus(0) + log(expm1(init)): negative for init = 0.3
    s = layer.scale_model().detach().numpy()
    np.testing.assert_allclose(s[:10], np.log(2.0) + np.log(np.expm1(0.3)), rtol=1e-6)
    np.testing.assert_allclose(s[10:], np.log(2.0) + np.log(np.expm1(0.7)), rtol=1e-6)
    assert s[0] < 0 < s[-1]
    layer2 = GaussianKernelsLayer(n_centers=10, n_dims=2)
    layer2.set_center_points(np.linspace(-1, 1, 200).reshape((100, 2)))
    assert layer2(torch.ones((3, 20))).sample().shape == (3, 2)


def test_cpu_tensors_are_rejected_not_computed():
    dist = InverseNormalizingFlowLayer(["radial"], 1, False)(torch.zeros((4, 3)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        dist.log_prob(torch.zeros((4, 1)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        FLOWS["planar"](torch.ones((2, 3)), 1).forward(torch.zeros((2, 1)))
