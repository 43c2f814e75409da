"""Input pipelines in front of the step (pipeline.py): the device-resident loader visits every sequence exactly once
per epoch and feeds train_loop unchanged; the host prefetcher hands out exactly the submitted batches while copies
overlap compute."""
import numpy as np
import pytest
import torch

import groove_oracle as G
from _util import build_model
from transformergrooveinfilling_b200 import FusedSGD, calculate_loss, train_loop
from transformergrooveinfilling_b200.pipeline import DeviceResidentLoader, HostBatchPrefetcher, HostPredictor

pytestmark = pytest.mark.gpu


def test_device_resident_loader_is_a_permutation_per_epoch():
    s = 1000
    x = torch.arange(s, dtype=torch.float32).view(s, 1, 1).expand(s, 32, 16).contiguous()
    y = (torch.arange(s, dtype=torch.float32) * 2).view(s, 1, 1).expand(s, 32, 27).contiguous()
    ld = DeviceResidentLoader(x, y, 96, "cuda", shuffle=True, seed=3)
    assert len(ld) == 11 and len(ld.dataset) == s
    orders = []
    for _ in range(2):
        seen = []
        for bx, by, idx in ld:
            assert bx.is_cuda and bx.shape[1:] == (32, 16) and by.shape[1:] == (32, 27)
            assert torch.equal(bx[:, 0, 0], idx.float()) and torch.equal(by[:, 5, 7], 2 * idx.float())
            seen.append(idx.cpu())
        seen = torch.cat(seen)
        assert seen.numel() == s and torch.equal(torch.sort(seen).values, torch.arange(s))
        orders.append(seen)
    assert not torch.equal(orders[0], orders[1])            # a fresh permutation every epoch
    tail = [b[0].shape[0] for b in DeviceResidentLoader(x, y, 96, "cuda", shuffle=False, drop_last=True)]
    assert tail == [96] * 10


def test_train_loop_accepts_the_device_resident_loader():
    cfg = G.GrooveCfg(32, 4, 16, 2, 0, 16, 27, dropout=0.0)
    m, P = build_model(cfg)
    x, y = G.det_batch(cfg, 64)
    opt = FusedSGD(m, 0.05)
    ld = DeviceResidentLoader(x, y, 64, "cuda", shuffle=False)
    loss = train_loop(ld, m, calculate_loss, None, None, opt, 0, False, "cuda", True, hit_loss_penalty=0.47)
    loss6, _, _ = G.train_step_oracle(P, cfg, x, y, 0.47, G.DropCtx(0.0, 0, 0, 0, False))
    assert abs(loss - loss6[0]) / abs(loss6[0]) < 1e-4


def test_host_prefetcher_hands_out_submitted_batches_in_order():
    n = 4096
    dev = torch.device("cuda")
    f = HostBatchPrefetcher(dev, (n, 32, 16), (n, 32, 27))
    hx = [torch.full((n, 32, 16), float(i)).pin_memory() for i in range(5)]
    hy = [torch.full((n, 32, 27), float(-i)).pin_memory() for i in range(5)]
    sums = []
    f.submit(hx[0], hy[0])
    for i in range(5):
        xd, yd = f.get()
        if i + 1 < 5:
            f.submit(hx[i + 1], hy[i + 1])
        big = torch.randn(2048, 2048, device=dev)
        for _ in range(4):
            big = big @ big * 1e-3                            # keep the compute stream busy while the next copy runs
        sums.append((xd.mean() + 0 * big[0, 0].nan_to_num(), yd.mean()))
    with pytest.raises(RuntimeError):
        f.get()
    torch.cuda.synchronize()
    for i, (sx, sy) in enumerate(sums):
        assert float(sx) == float(i) and float(sy) == float(-i)
    assert f.h2d_bytes == 5 * n * 32 * (16 + 27) * 4


@pytest.mark.parametrize("n_dec,precision", [(0, "fp32"), (0, "bf16"), (2, "bf16")])
def test_host_predictor_equals_predict_on_the_whole_array(n_dec, precision):
    """HostPredictor: chunked, copy-overlapped predict() over a host array returns the [N, 32, 27] array the evaluator builds
    from model.predict (evaluator.py:171-175): hits as 0 / 1, velocities, offsets — ragged last chunk, pinned or pageable input,
    caller-provided output, repeated calls on the same slots."""
    cfg = G.GrooveCfg(32, 4, 64, 2, n_dec, 16, 27, dropout=0.1)
    m, _ = build_model(cfg, precision=precision)
    x, _y = G.det_batch(cfg, 1000)
    x = torch.as_tensor(x, dtype=torch.float32)
    h, v, o = m.predict(x.cuda())
    want = torch.cat((h.float(), v, o), 2).cpu()
    hp = HostPredictor(m, chunk=384)                       # 1000 = 2 x 384 + 232: both slots reused, ragged tail
    got = hp.predict(x.pin_memory())
    assert got.shape == (1000, 32, 27) and got.is_pinned() and not m.training
    assert torch.equal(got, want)                          # same kernels on the same rows: sequences are independent
    out = torch.empty(1000, 32, 27).pin_memory()
    assert hp.predict(x, out=out) is out and torch.equal(out, want)        # pageable input, second call
    assert hp.h2d_bytes == 2 * x.numel() * 4 and hp.d2h_bytes == 2 * want.numel() * 4
    with pytest.raises(ValueError):
        hp.predict(x.cuda())
    with pytest.raises(ValueError):
        hp.predict(x[:, :, :8])
