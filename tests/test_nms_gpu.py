"""Bit-exact parity of the device-side multi-label NMS against the oracle restatement of csrc/cuda/ml_nms.cu."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _boxes(seed, n, nlabels, img=(1344, 800), cluster=True):
    g = torch.Generator().manual_seed(seed)
    if cluster:  # many heavily overlapping boxes around a few centres, like dense ATSS candidates
        nc = max(1, n // 40)
        cx = torch.rand(nc, generator=g) * img[0]
        cy = torch.rand(nc, generator=g) * img[1]
        which = torch.randint(0, nc, (n,), generator=g)
        x = cx[which] + torch.randn(n, generator=g) * 12
        y = cy[which] + torch.randn(n, generator=g) * 12
    else:
        x = torch.rand(n, generator=g) * img[0]
        y = torch.rand(n, generator=g) * img[1]
    w = torch.rand(n, generator=g) * 150 + 10
    h = torch.rand(n, generator=g) * 150 + 10
    boxes = torch.stack([x - w / 2, y - h / 2, x + w / 2, y + h / 2], 1)
    scores = torch.rand(n, generator=g)
    scores[::7] = scores[0]  # exact score ties
    labels = torch.randint(1, nlabels + 1, (n,), generator=g).float()
    return boxes, scores, labels


@pytest.mark.parametrize("n,nlabels", [(1, 1), (63, 2), (64, 1), (65, 3), (1000, 10), (5000, 80), (4097, 1)])
def test_ml_nms_bit_exact(dev, n, nlabels):
    from mqdet_b200 import ops
    from oracle import restate
    boxes, scores, labels = _boxes(n + nlabels, n, nlabels)
    ref = restate.ml_nms(boxes, scores, labels, 0.6)
    keep = ops.ml_nms(boxes.to(dev), scores.to(dev), labels.to(dev), 0.6)
    assert keep.dtype == torch.int64
    assert torch.equal(keep.cpu(), ref), f"kept sets differ: {keep.numel()} vs {ref.numel()}"


def test_argsort_matches_stable_sort(dev):
    from mqdet_b200 import ops
    g = torch.Generator().manual_seed(1)
    s = torch.rand(5000, generator=g)
    s[100:200] = 0.5
    s[7] = -1.0
    order = ops.argsort_desc(s.to(dev)).cpu()
    ref = torch.sort(s, descending=True, stable=True)[1]
    assert torch.equal(order, ref)


def test_empty_and_topk_cut(dev):
    from mqdet_b200 import ops
    from oracle import restate
    e = ops.ml_nms(torch.zeros(0, 4, device=dev), torch.zeros(0, device=dev), torch.zeros(0, device=dev), 0.6)
    assert e.numel() == 0 and e.dtype == torch.int64  # reference: empty tensor (csrc/ml_nms.h:19-20)
    boxes, scores, labels = _boxes(99, 3000, 20, cluster=False)
    ref = restate.select_over_all_levels(boxes, scores, labels, 0.6, 100)
    keep, num = ops.ml_nms_device(boxes.to(dev), scores.to(dev), labels.to(dev), 0.6, max_det=100)
    got = keep[: int(num.item())].cpu()
    assert torch.equal(got, ref)
    assert got.numel() >= 100
