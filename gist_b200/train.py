"""Training-step helpers shared by the trainers and bench.py (the L4 glue of
cluster_gcn_ist_distrib.py:370-479, kept thin)."""
import torch
import torch.nn.functional as F  # noqa: F401

from . import ops
from .optim import Adam


def masked_cross_entropy(pred, labels, mask):
    """CrossEntropyLoss()(pred[mask], labels[mask]) (…distrib.py:413-414) without the
    boolean-index host sync: mean of the per-row losses over the masked rows, in the fused
    CE kernels (csrc/fused.cu)."""
    return ops.masked_cross_entropy(pred, labels, mask.bool())


def loss_and_backward(pred, labels, mask, loss_out=None):
    """``loss = CrossEntropyLoss()(pred[mask], labels[mask]); loss.backward()`` (…distrib.py:413-415)
    with the loss and its gradient seed d loss / d pred produced by one kernel: the backward pass
    starts from ``pred`` directly.  Returns the detached 0-d loss (``loss_out``: a persistent 2-float
    buffer the kernel writes the loss into; the returned tensor is then its element 0)."""
    if pred.shape[0] == 0:
        loss = masked_cross_entropy(pred, labels, mask)
        loss.backward()
        if loss_out is not None:
            loss_out[0].copy_(loss.detach())
            return loss_out[0]
        return loss.detach()
    loss, dpred = ops.masked_ce_loss_and_grad(pred, labels, mask.bool(), out=loss_out)
    pred.backward(dpred)
    return loss


def make_optimizer(params, lr, weight_decay):
    """Adam as the reference builds it (…distrib.py:405-407: torch.optim.Adam(lr, weight_decay)),
    with the update of all tensors in one launch."""
    return Adam(params, lr=lr, weight_decay=weight_decay)


def train_step(model, optimizer, cluster):
    """zero_grad -> forward -> CE on train rows -> backward -> Adam (…distrib.py:408-417).
    Returns the loss as a 0-d device tensor (no sync)."""
    optimizer.zero_grad(set_to_none=True)
    pred = model(cluster)
    loss = loss_and_backward(pred, cluster.ndata['label'], cluster.ndata['train_mask'])
    optimizer.step()
    return loss


@torch.no_grad()
def evaluate(model, g, labels, mask, method='acc'):
    """cluster_gcn/utils.py:70-80: full-graph inference; accuracy (== micro-F1 for
    single-label prediction) of argmax over the masked rows."""
    assert method in ['acc', 'f1'], 'invalid method'
    model.eval()
    logits = model(g)
    pred = logits.argmax(dim=1)
    m = mask.bool()
    total = int(m.sum().item())
    if total == 0:
        return -1
    return float(((pred == labels) & m).sum().item()) / total


@torch.no_grad()
def evaluate_masks(model, g, labels, masks, method='acc'):
    """The reference evaluates val and test with TWO full-graph forward passes per eval point
    (cluster_gcn_ist_distrib.py:439-446 -> utils.py:70-80).  One pass serves any number of masks:
    returns [accuracy over mask_i].  Counts stay on the device until one final readback."""
    assert method in ['acc', 'f1'], 'invalid method'
    model.eval()
    pred = model(g).argmax(dim=1)
    hit = pred == labels
    M = torch.stack([m.bool() for m in masks])                        # [k, n]
    cnt = torch.stack(((M & hit).sum(1), M.sum(1))).cpu()             # one D2H
    return [float(cnt[0, i]) / int(cnt[1, i]) if int(cnt[1, i]) else -1 for i in range(len(masks))]
