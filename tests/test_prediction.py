"""Host logic of the device-resident tiled predictor (torch-em_b200/util/prediction.py) against the CPU restatement of the
reference's predict_with_halo (oracle/prediction.py), with a small torch function standing in for the network (the real
U-Net refuses CPU tensors by design).  Mirrors test/util/test_prediction.py:20-31 (coverage / shape) and adds values."""
import numpy as np
import pytest
import torch

from oracle import prediction as opred
from torch_em_b200.util import Blocking, predict_with_halo, standardize


class TinyNet(torch.nn.Module):
    """2 output channels from a 3x3x3 box filter and a pointwise map: depends on the halo, cheap, deterministic."""
    out_channels = 2

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.linspace(-1, 1, 27).reshape(1, 1, 3, 3, 3), requires_grad=False)

    def forward(self, x):
        x0 = x[:, :1]
        a = torch.nn.functional.conv3d(x0, self.w, padding=1)
        return torch.cat([a, torch.tanh(x0)], dim=1)


def np_net(net):
    return lambda a: net(torch.from_numpy(a)).numpy()


def test_blocking_covers_volume_once():
    shape, bs = (37, 20, 45), (16, 16, 16)
    blk = Blocking([0, 0, 0], list(shape), list(bs))
    assert blk.number_of_blocks == 3 * 2 * 3
    seen = np.zeros(shape, dtype=int)
    for i in range(blk.number_of_blocks):
        b = blk.get_block(i)
        assert all(s <= q for s, q in zip(b.shape, bs))
        seen[tuple(slice(x, y) for x, y in zip(b.begin, b.end))] += 1
    assert (seen == 1).all()
    # C order: the last axis runs fastest
    assert blk.get_block(1).begin == [0, 0, 16] and blk.get_block(3).begin == [0, 16, 0]
    with pytest.raises(IndexError):
        blk.get_block(blk.number_of_blocks)


def test_standardize_matches_reference_formula():
    x = np.random.default_rng(0).random((9, 10, 11)).astype("float32") * 7 + 3
    np.testing.assert_allclose(standardize(torch.from_numpy(x)).numpy(), opred.standardize(x), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shape,block_shape,halo", [((24, 20, 28), (16, 16, 16), (4, 4, 4)), ((16, 16, 16), (8, 8, 8), (3, 5, 2)),
                                                    ((10, 33, 17), (8, 16, 16), (2, 6, 6))])
def test_predict_with_halo_matches_oracle(shape, block_shape, halo):
    rng = np.random.default_rng(1)
    vol = rng.random(shape).astype("float32")
    net = TinyNet()
    ref = opred.predict_with_halo(vol, np_net(net), block_shape, halo, n_out=2)
    out = predict_with_halo(vol, net, ["cpu"], block_shape, halo)
    assert out.shape == (2,) + shape and out.dtype == np.float32
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5)
    # with a channel axis and a caller-provided output
    buf = np.zeros((2,) + shape, dtype="float32")
    ret = predict_with_halo(vol[None], net, ["cpu"], block_shape, halo, output=buf, with_channels=True)
    assert ret is buf
    ref_c = opred.predict_with_halo(vol[None], np_net(net), block_shape, halo, n_out=2, with_channels=True)
    np.testing.assert_allclose(buf, ref_c, rtol=1e-4, atol=1e-5)


def test_unsupported_arguments_raise():
    vol = np.zeros((8, 8, 8), dtype="float32")
    with pytest.raises(NotImplementedError):
        predict_with_halo(vol, TinyNet(), ["cpu"], (8, 8, 8), (2, 2, 2), mask=np.ones((8, 8, 8)))
    with pytest.raises(ValueError):
        predict_with_halo(vol, TinyNet(), ["cpu"], (8, 8), (2, 2, 2))


@pytest.mark.gpu
def test_predict_with_halo_unet_on_gpu():
    """The real U-Net (fp32, exact CUDA-core path) behind the device-resident predictor vs the CPU restatement of the
    reference's predict_with_halo running the oracle U-Net on the same weights."""
    import torch_em_b200 as tb
    from oracle import unet as ounet
    torch.manual_seed(0)
    net = tb.UNet3d(1, 2, depth=2, initial_features=8, final_activation="Sigmoid").to("cuda:0")
    vol = np.random.default_rng(2).random((24, 32, 40)).astype("float32")
    block_shape, halo = (16, 16, 16), (8, 8, 8)
    out = predict_with_halo(vol, net, [0], block_shape, halo)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}

    def cpu_net(a):
        with torch.no_grad():
            return ounet.unet3d_forward(torch.from_numpy(a), sd, [2, 2], final_activation="Sigmoid").numpy()

    ref = opred.predict_with_halo(vol, cpu_net, block_shape, halo, n_out=2)
    assert out.shape == ref.shape
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)
