"""Fused masked cross-entropy: the caller-side packing loop + nn.CrossEntropyLoss of
run_gun.py:189-197 (B Python slices + torch.cat of ~70 MB + log-softmax) as ONE kernel that reads the
(B,L,V) logits once and writes d(loss)/d(logits) in the same pass (no packing copy)."""
import torch

from . import ops


class _PackedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, lens_t, inv_count, inv_dev=None, unit_grad=False):
        logits = logits if logits.is_contiguous() else logits.contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        ops.backend().ce_masked(logits, targets.contiguous(), lens_t, loss, dlogits, inv_count, inv_dev)
        ctx.save_for_backward(dlogits)
        ctx.unit_grad = unit_grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        if ctx.unit_grad:                 # caller guarantees d(total)/d(loss) == 1: no 70 MB scaling pass
            return dlogits, None, None, None, None, None
        return dlogits * g, None, None, None, None, None


def packed_cross_entropy(outputs, targets, cap_lens, inv_count_dev=None, unit_grad=False):
    """outputs (B,L,V) raw logits, targets (B,L) int64, cap_lens: python ints (or a device int32 tensor together with
    inv_count_dev = 1/sum(lens) as a 1-element device tensor: graph-capturable form).  Mean NLL over the tokens.
    unit_grad=True: the caller promises the loss enters the differentiated total with coefficient exactly 1
    (loss.backward(), or cap_loss + lambda * other: run_gun.py:198,231), so backward hands out the gradient computed
    by the forward kernel without a scaling pass."""
    if torch.is_tensor(cap_lens):
        return _PackedCE.apply(outputs, targets[:, :outputs.shape[1]], cap_lens, 0.0, inv_count_dev, unit_grad)
    lens_t = torch.as_tensor(list(cap_lens), dtype=torch.int32).to(outputs.device, non_blocking=True)
    L = outputs.shape[1]
    n = sum(min(int(c), L) for c in cap_lens)
    return _PackedCE.apply(outputs, targets[:, :L], lens_t, 1.0 / max(n, 1), None, unit_grad)
