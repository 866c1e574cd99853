"""Host logic of the device-resident tiled predictor (torch-em_b200/util/prediction.py) against the CPU restatement of the
reference's predict_with_halo (oracle/prediction.py), with a small torch function standing in for the network (the real
U-Net refuses CPU tensors by design).  Mirrors test/util/test_prediction.py:20-31 (coverage / shape) and adds values."""
import numpy as np
import pytest
import torch

from oracle import prediction as opred
from torch_em_b200.util import Blocking, predict_with_halo, predict_with_halo_pipelined, standardize


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


def _all_argument_cases(rng, shape):
    """(name, kwargs for ours, kwargs for the oracle) covering every argument of prediction.py:145-164."""
    mask = np.zeros(shape, dtype="uint8")
    mask[2:-3, :shape[1] // 2, 3:] = 1
    mask[rng.random(shape) < 0.2] = 0
    return [
        ("mask", dict(mask=mask), dict(mask=mask)),
        ("roi", dict(roi=(slice(4, 20), slice(None), slice(8, None))), dict(roi=(slice(4, 20), slice(None), slice(8, None)))),
        ("iter_list", dict(iter_list=[0, 3, 5]), dict(iter_list=[0, 3, 5])),
        ("grid_shift", dict(grid_shift=(0.0, 0.25, 0.5)), dict(grid_shift=(0.0, 0.25, 0.5))),
        ("grid_shift+mask", dict(grid_shift=(0.5, 0.0, 0.25), mask=mask), dict(grid_shift=(0.5, 0.0, 0.25), mask=mask)),
        ("no preprocess", dict(preprocess=None), dict(preprocess=None)),
    ]


def check_all_arguments(device_ids, net, to_np_net):
    rng = np.random.default_rng(4)
    shape, block_shape, halo = (24, 20, 28), (16, 8, 16), (4, 4, 4)
    vol = rng.random(shape).astype("float32")
    for name, ours_kw, ref_kw in _all_argument_cases(rng, shape):
        ref = opred.predict_with_halo(vol, to_np_net, block_shape, halo, n_out=2, **ref_kw)
        out = predict_with_halo(vol, net, device_ids, block_shape, halo, **ours_kw)
        assert out.shape == ref.shape, name
        np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5, err_msg=name)
    # list of (array, channel slice) outputs, one of them without a channel axis; pre-filled outputs keep what is not written
    o1, o2 = np.full((1,) + shape, 7.0, dtype="float32"), np.full(shape, 7.0, dtype="float32")
    predict_with_halo(vol, net, device_ids, block_shape, halo, output=[(o1, np.s_[0:1]), (o2, 1)], iter_list=[1, 2])
    r1, r2 = np.full((1,) + shape, 7.0, dtype="float32"), np.full(shape, 7.0, dtype="float32")
    opred.predict_with_halo(vol, to_np_net, block_shape, halo, n_out=2, output=[(r1, np.s_[0:1]), (r2, 1)], iter_list=[1, 2])
    np.testing.assert_allclose(o1, r1, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(o2, r2, rtol=1e-4, atol=1e-5)
    # host callbacks (numpy in, numpy out) take the streaming path with the same results
    skip = lambda a: bool(a.mean() > 0.505)                                   # noqa: E731
    post = lambda p: p[:1] * 2.0                                              # noqa: E731
    ref = opred.predict_with_halo(vol, to_np_net, block_shape, halo, n_out=1, skip_block=skip, postprocess=post,
                                  output=np.zeros((1,) + shape, dtype="float32"))
    out = predict_with_halo(vol, net, device_ids, block_shape, halo, skip_block=skip, postprocess=post,
                            output=np.zeros((1,) + shape, dtype="float32"))
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5)
    # pipelined variant: same results, batches of blocks, grid_shift refused like the reference (prediction.py:555-559)
    ref = opred.predict_with_halo(vol, to_np_net, block_shape, halo, n_out=2)
    out = predict_with_halo_pipelined(vol, net, device_ids, block_shape, halo, batch_size=3, num_prefetch_workers=2)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5)
    with pytest.raises(NotImplementedError, match="grid_shift"):
        predict_with_halo_pipelined(vol, net, device_ids, block_shape, halo, grid_shift=(0, 0, 0.5))
    with pytest.raises(ValueError, match="grid_shift"):
        predict_with_halo(vol, net, device_ids, block_shape, halo, grid_shift=(0, 0, 0.5), output=np.zeros((2,) + shape, "float32"))


def test_every_argument_matches_oracle_streaming_path():
    net = TinyNet()
    check_all_arguments(["cpu"], net, np_net(net))


def test_argument_validation():
    vol = np.zeros((8, 8, 8), dtype="float32")
    with pytest.raises(ValueError):
        predict_with_halo(vol, TinyNet(), ["cpu"], (8, 8), (2, 2, 2))
    with pytest.raises(ValueError):
        predict_with_halo(vol, TinyNet(), [], (8, 8, 8), (2, 2, 2))


@pytest.mark.gpu
def test_every_argument_matches_oracle_device_path():
    """The device-resident path (gather / standardize / scatter kernels of csrc/tiling.cu) with a plain torch model on the GPU."""
    import torch_em_b200 as tb
    net = TinyNet().to("cuda:0")
    cpu = TinyNet()
    tb.reset_launch_count()
    check_all_arguments([0], net, np_net(cpu))
    assert tb.launch_count() > 20, "the tiling kernels did not run"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["uint8", "uint16", "int16", "float64", "int64"])
def test_device_path_raw_dtypes(dtype):
    """EM volumes are usually uint8 / uint16: the gather kernel converts on the fly (torch has no uint16 index_select)."""
    rng = np.random.default_rng(5)
    vol = (rng.random((20, 24, 18)) * 200).astype(dtype)
    net, cpu = TinyNet().to("cuda:0"), TinyNet()
    ref = opred.predict_with_halo(vol, np_net(cpu), (8, 16, 16), (3, 2, 4), n_out=2)
    out = predict_with_halo(vol, net, [0], (8, 16, 16), (3, 2, 4))
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("overlap", ["1", "0"])
@pytest.mark.parametrize("order", ["ascending", "descending", "shuffled", "partial"])
def test_device_path_transfer_overlap(monkeypatch, overlap, order):
    """The volume goes up in row ranges ahead of the batches that read them and finished output rows go down behind them
    (side streams).  Whatever the block order (iter_list) -- ascending overlaps, anything else degrades to up-front / at-the-end
    copies -- the result is the reference's; several row slabs, truncated last blocks, a mask and a batch of 3."""
    monkeypatch.setenv("B200EM_PREDICT_OVERLAP", overlap)
    rng = np.random.default_rng(11)
    vol = (rng.random((50, 22, 26)) * 100).astype("float32")
    mask = np.ones(vol.shape, dtype=bool)
    mask[18:31, :, 5:20] = False
    mask[16:24, 0:16, 0:16] = False                       # block 8 has an empty inner mask: skipped
    net, cpu = TinyNet().to("cuda:0"), TinyNet()
    block_shape, halo = (8, 16, 16), (3, 2, 4)
    n_blocks = 7 * 2 * 2
    ids = {"ascending": None, "descending": list(range(n_blocks))[::-1], "shuffled": [int(i) for i in rng.permutation(n_blocks)],
           "partial": [3, 4, 17, 9, 27]}[order]
    ref = opred.predict_with_halo(vol, np_net(cpu), block_shape, halo, n_out=2, mask=mask, iter_list=ids, preprocess=opred.standardize)
    out = predict_with_halo_pipelined(vol, net, [0], block_shape, halo, mask=mask, iter_list=ids, preprocess=standardize, batch_size=3)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)
    from torch_em_b200.util import prediction as P
    assert P.last_timing["overlap"] == (overlap == "1")


@pytest.mark.gpu
def test_device_path_two_devices():
    """``gpu_ids=[0, 1]``: block i runs on device i % 2 (prediction.py:249-250), one replica, one thread and one pair of transfer
    streams per device; the two partial results are merged on the host.  Skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(12)
    vol = (rng.random((40, 22, 26)) * 100).astype("float32")
    net, cpu = TinyNet().to("cuda:0"), TinyNet()
    block_shape, halo = (8, 16, 16), (3, 2, 4)
    ref = opred.predict_with_halo(vol, np_net(cpu), block_shape, halo, n_out=2)
    out = predict_with_halo(vol, net, [0, 1], block_shape, halo)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)
    out = predict_with_halo_pipelined(vol, net, ["cuda:1", "cuda:0"], block_shape, halo, batch_size=2)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_device_path_2d_and_channels():
    class Net2d(torch.nn.Module):
        out_channels = 3

        def __init__(self):
            super().__init__()
            self.dummy = torch.nn.Parameter(torch.zeros(1))       # predict_with_halo reads the device from the parameters

        def forward(self, x):
            return torch.cat([x.mean(1, keepdim=True), torch.tanh(x[:, :1]), torch.nn.functional.avg_pool2d(x[:, 1:2], 3, 1, 1)], 1)

    rng = np.random.default_rng(6)
    vol = rng.random((2, 40, 52)).astype("float32")
    net = Net2d()
    ref = opred.predict_with_halo(vol, np_net(net), (16, 32), (4, 6), n_out=3, with_channels=True)
    out = predict_with_halo(vol, net.to("cuda:0"), ["cuda:0"], (16, 32), (4, 6), with_channels=True)
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5)


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
