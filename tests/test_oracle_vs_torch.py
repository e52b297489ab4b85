"""Cross-check of the restated ops that the reference has neither CPU code nor tests for
(SURVEY.md 8c "parity unpinned") against torch CPU ops/autograd in float64."""
import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

from oracle import pyoracle as orc

rng = np.random.default_rng(7)

CONV_CASES = [
    # N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw
    (2, 3, 5, 6, 8, 0, 0, 1, 1, 3, 5),
    (2, 3, 5, 6, 7, 3, 2, 3, 2, 3, 5),
    (3, 4, 6, 9, 9, 1, 1, 1, 1, 3, 3),
    (2, 2, 4, 11, 10, 2, 2, 2, 2, 5, 5),
    (1, 3, 8, 23, 23, 0, 0, 4, 4, 11, 11),   # AlexNet conv1 geometry, (23-11)%4 == 0
    (2, 3, 4, 12, 13, 1, 0, 2, 3, 4, 3),     # (H+2p-f)%s != 0 -> floor
]


def _t(a, *shape):
    return torch.tensor(np.asarray(a, np.float64).reshape(shape), requires_grad=True)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_all_directions(case):
    N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw = case
    x = rng.normal(0, 1, N * Ci * H * W).astype(np.float32)
    w = rng.normal(0, 1, Co * Ci * fh * fw).astype(np.float32)
    b = rng.normal(0, 1, Co).astype(np.float32)
    tx, tw, tb = _t(x, N, Ci, H, W), _t(w, Co, Ci, fh, fw), _t(b, Co)
    ty = Fn.conv2d(tx, torch.flip(tw, (2, 3)), tb, stride=(sv, sh), padding=(ph, pw))
    y = orc.conv_forward(x, w, b, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw)
    assert y.size == ty.numel()
    np.testing.assert_allclose(y, ty.detach().numpy().ravel(), rtol=1e-4, atol=1e-4)
    dy = rng.normal(0, 1, y.size).astype(np.float32)
    ty.backward(torch.tensor(dy.astype(np.float64).reshape(ty.shape)))
    np.testing.assert_allclose(orc.conv_backward_data(dy, w, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw),
                               tx.grad.numpy().ravel(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(orc.conv_backward_filter(x, dy, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw),
                               tw.grad.numpy().ravel(), rtol=1e-4, atol=2e-4)
    _, _, Ho, Wo = ty.shape
    np.testing.assert_allclose(orc.conv_backward_bias(dy, N, Co, Ho, Wo), tb.grad.numpy().ravel(),
                               rtol=1e-5, atol=1e-5)


POOL_CASES = [
    # N, C, H, W, sv, sh, wh, ww, ph, pw
    (2, 3, 4, 4, 1, 1, 3, 3, 0, 0),
    (2, 3, 4, 4, 2, 2, 3, 3, 1, 1),
    (1, 2, 4, 4, 3, 3, 4, 4, 2, 2),
    (2, 4, 13, 13, 2, 2, 3, 3, 0, 0),   # AlexNet pool5
    (2, 2, 12, 12, 3, 3, 3, 3, 0, 0),   # LeNet pool2
    (1, 3, 14, 14, 2, 2, 3, 3, 0, 0),   # GoogLeNet overhang: (14-3)%2 != 0
    (1, 3, 7, 9, 1, 1, 3, 3, 1, 1),     # GoogLeNet inception pool 3x3/1 pad 1
]


def _torch_pool_sizes(H, W, sv, sh, wh, ww, ph, pw):
    return orc.pooled_size(H, ph, wh, sv), orc.pooled_size(W, pw, ww, sh)


@pytest.mark.parametrize("case", POOL_CASES)
def test_max_pooling(case):
    N, C, H, W, sv, sh, wh, ww, ph, pw = case
    x = rng.normal(0, 1, N * C * H * W).astype(np.float32)
    tx = _t(x, N, C, H, W)
    ty = Fn.max_pool2d(tx, (wh, ww), (sv, sh), (ph, pw), ceil_mode=True)
    Ho, Wo = _torch_pool_sizes(H, W, sv, sh, wh, ww, ph, pw)
    assert tuple(ty.shape[2:]) == (Ho, Wo)
    y = orc.max_pooling_forward(x, N, C, H, W, sv, sh, wh, ww, ph, pw)
    np.testing.assert_array_equal(y, ty.detach().numpy().astype(np.float32).ravel())
    dy = rng.normal(0, 1, y.size).astype(np.float32)
    ty.backward(torch.tensor(dy.astype(np.float64).reshape(ty.shape)))
    # no ties in continuous random data -> argmax routing is unambiguous
    np.testing.assert_allclose(orc.max_pooling_backward(x, y, dy, N, C, H, W, sv, sh, wh, ww, ph, pw),
                               tx.grad.numpy().ravel(), rtol=1e-6, atol=1e-6)


def test_max_pooling_backward_tie_rule():
    """All-equal window (post-ReLU zeros): the first position in (h-major, w-minor) scan gets dy."""
    x = np.zeros(16, np.float32)
    y = orc.max_pooling_forward(x, 1, 1, 4, 4, 2, 2, 2, 2, 0, 0)
    dy = np.array([1, 2, 3, 4], np.float32)
    dx = orc.max_pooling_backward(x, y, dy, 1, 1, 4, 4, 2, 2, 2, 2, 0, 0).reshape(4, 4)
    want = np.zeros((4, 4), np.float32)
    want[0, 0], want[0, 2], want[2, 0], want[2, 2] = 1, 2, 3, 4
    np.testing.assert_array_equal(dx, want)


@pytest.mark.parametrize("case", POOL_CASES)
def test_average_pooling(case):
    N, C, H, W, sv, sh, wh, ww, ph, pw = case
    Ho, Wo = _torch_pool_sizes(H, W, sv, sh, wh, ww, ph, pw)
    x = rng.normal(0, 1, N * C * H * W).astype(np.float32)
    # COUNT_INCLUDE_PADDING with a fixed wh*ww divisor == sum-pool over the zero-padded image / (wh*ww),
    # windows allowed to overhang bottom/right (torch's own divisor clips there, so build it by hand)
    need_h, need_w = (Ho - 1) * sv + wh, (Wo - 1) * sh + ww
    tx = _t(x, N, C, H, W)
    xp = Fn.pad(tx, (pw, max(0, need_w - W - pw), ph, max(0, need_h - H - ph)))
    ty = Fn.avg_pool2d(xp, (wh, ww), (sv, sh), 0, ceil_mode=False, count_include_pad=True)[:, :, :Ho, :Wo]
    y = orc.average_pooling_forward(x, N, C, H, W, sv, sh, wh, ww, ph, pw)
    np.testing.assert_allclose(y, ty.detach().numpy().ravel(), rtol=1e-5, atol=1e-6)
    dy = rng.normal(0, 1, y.size).astype(np.float32)
    ty.backward(torch.tensor(dy.astype(np.float64).reshape(ty.shape)))
    np.testing.assert_allclose(orc.average_pooling_backward(x, y, dy, N, C, H, W, sv, sh, wh, ww, ph, pw),
                               tx.grad.numpy().ravel(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("N,C,H,W,size", [(2, 7, 3, 4, 5), (1, 16, 5, 5, 5), (2, 5, 2, 3, 3), (1, 6, 2, 2, 4)])
def test_lrn(N, C, H, W, size):
    alpha, beta = 1e-2, 0.75
    x = rng.normal(0, 2, N * C * H * W).astype(np.float32)
    tx = _t(x, N, C, H, W)
    if size % 2:  # torch's LRN matches Caffe's window only for odd sizes
        ty = Fn.local_response_norm(tx, size, alpha=alpha, beta=beta, k=1.0)
    else:  # even size: Caffe window is [c - (size-1)//2, c + size//2]
        pre, post = (size - 1) // 2, size - (size - 1) // 2 - 1
        sq = Fn.pad(tx * tx, (0, 0, 0, 0, pre, post))
        ssum = sum(sq[:, i:i + C] for i in range(size))
        ty = tx * (1.0 + alpha / size * ssum) ** (-beta)
    y, scale = orc.lrn_forward(x, size, alpha, beta, N, C, W, H)
    np.testing.assert_allclose(y, ty.detach().numpy().ravel(), rtol=2e-5, atol=1e-6)
    dy = rng.normal(0, 1, y.size).astype(np.float32)
    ty.backward(torch.tensor(dy.astype(np.float64).reshape(ty.shape)))
    dx = orc.lrn_backward(x, y, scale, dy, size, alpha, beta, N, C, W, H)
    np.testing.assert_allclose(dx, tx.grad.numpy().ravel(), rtol=1e-4, atol=1e-5)


def test_softmax_modes_and_backward():
    N, C, H, W = 3, 5, 2, 4
    x = rng.normal(0, 3, N * C * H * W).astype(np.float32)
    dy = rng.normal(0, 1, x.size).astype(np.float32)
    tx = _t(x, N, C * H * W)
    ty = torch.softmax(tx, 1)
    y = orc.instance_softmax_forward(x, N, C, H, W)
    np.testing.assert_allclose(y, ty.detach().numpy().ravel(), rtol=1e-5, atol=1e-7)
    ty.backward(torch.tensor(dy.astype(np.float64).reshape(ty.shape)))
    np.testing.assert_allclose(orc.instance_softmax_backward(dy, y, N, C, H, W), tx.grad.numpy().ravel(),
                               rtol=1e-4, atol=1e-6)
    tx = _t(x, N, C, H, W)
    ty = torch.softmax(tx, 1)
    y = orc.channel_softmax_forward(x, N, C, H, W)
    np.testing.assert_allclose(y, ty.detach().numpy().ravel(), rtol=1e-5, atol=1e-7)
    ty.backward(torch.tensor(dy.astype(np.float64).reshape(ty.shape)))
    np.testing.assert_allclose(orc.channel_softmax_backward(dy, y, N, C, H, W), tx.grad.numpy().ravel(),
                               rtol=1e-4, atol=1e-6)


def test_activation_backward():
    x = rng.normal(0, 1, 1000).astype(np.float32)
    dy = rng.normal(0, 1, 1000).astype(np.float32)
    for fwd, bwd, tf in ((orc.sigmoid_forward, orc.sigmoid_backward, torch.sigmoid),
                         (orc.relu_forward, orc.relu_backward, torch.relu),
                         (orc.tanh_forward, orc.tanh_backward, torch.tanh)):
        tx = _t(x, 1000)
        ty = tf(tx)
        ty.backward(torch.tensor(dy.astype(np.float64)))
        y = fwd(x)
        np.testing.assert_allclose(bwd(x, y, dy), tx.grad.numpy(), rtol=1e-5, atol=1e-6)
    # relu edge cases: -0.0 -> +0.0, NaN -> 0 (basic.cpp:430)
    e = np.array([-0.0, np.nan, 0.0, -1.0, 2.0], np.float32)
    np.testing.assert_array_equal(orc.relu_forward(e).view(np.uint32),
                                  np.array([0.0, 0.0, 0.0, 0.0, 2.0], np.float32).view(np.uint32))


def test_concat_slice_transpose_semantics():
    # concat two {W,H,C,N} batches on C (second-to-last dim): per-image block copies (cuda.cpp:97-121)
    N, H, W, C1, C2 = 3, 2, 3, 2, 4
    a = rng.normal(0, 1, (N, C1, H, W)).astype(np.float32)
    b = rng.normal(0, 1, (N, C2, H, W)).astype(np.float32)
    out = np.zeros(N * (C1 + C2) * H * W, np.float32)
    orc.copy_strided(a.ravel(), out.size, C1 * H * W, N, C1 * H * W, (C1 + C2) * H * W, dst=out)
    orc.copy_strided(b.ravel(), out.size, C2 * H * W, N, C2 * H * W, (C1 + C2) * H * W, dst_off=C1 * H * W, dst=out)
    np.testing.assert_array_equal(out.reshape(N, C1 + C2, H, W), np.concatenate([a, b], 1))
    # slice channels [1,3) back out
    sl = orc.copy_strided(out, N * 2 * H * W, 2 * H * W, N, (C1 + C2) * H * W, 2 * H * W, src_off=1 * H * W)
    np.testing.assert_array_equal(sl.reshape(N, 2, H, W), np.concatenate([a, b], 1)[:, 1:3])
    m, n = 5, 3
    mat = rng.normal(0, 1, m * n).astype(np.float32)
    np.testing.assert_array_equal(orc.transpose(mat, m, n).reshape(m, n), mat.reshape(n, m).T)


def test_rng_distribution():
    p = 0.3
    mask = orc.rand_bernoulli(200000, 42, p)
    assert set(np.unique(mask)) <= {0.0, 1.0}
    assert abs(mask.mean() - p) < 5e-3
    z = orc.randn(200000, 42, 1.5, 0.5)
    assert abs(z.mean() - 1.5) < 5e-3 and abs(z.std() - 0.5) < 5e-3
    assert not np.array_equal(orc.rand_bernoulli(64, 1, 0.5), orc.rand_bernoulli(64, 2, 0.5))


def test_sgd_update_matches_chain():
    n = 1000
    w, d, g = (rng.normal(0, 1, n).astype(np.float32) for _ in range(3))
    mom, lr, wd, B = np.float32(0.9), np.float32(0.01), np.float32(5e-4), 256
    w2, d2 = orc.sgd_momentum_update(w, d, g, mom, lr / B, lr * wd)
    dd = mom * d - np.float32(lr / B) * g - np.float32(lr * wd) * w
    np.testing.assert_allclose(d2, dd, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(w2, w + dd, rtol=1e-6, atol=1e-7)
