"""Evaluation glue of the hot path's caller side: decoding a data-parallel shard of the evaluation set and
merging the captions of ALL ranks.

Replaces, for any world size, the reference's `gather_results` loop (evaluate.py:101-116: one `decode_tokens` call per
clip, each comparing token by token on the device) plus the rank merge hard-coded for exactly four ranks
(run_gun.py:267-276: `results_multi = [None for _ in range(4)]` ... `{**results_multi[0], ..., **results_multi[3]}`).
The arithmetic is unchanged: `net(frames, regions, None)` on this rank's batches (greedy or beam search, whatever
`update_beam_size` selected), ids to strings through the decoder's own vocabulary (`layer.py:464-477`).

* one device->host copy of the (B, steps) id table per batch (not one sync per token),
* the per-rank `OrderedDict`s are merged in rank order with later ranks overwriting earlier ones - exactly what the
  reference's dict-unpacking merge does for the duplicate clips a `DistributedSampler` pads the last batch with.
"""
import collections

import torch
import torch.distributed as dist


def tokens_to_captions(decoder, token_table):
    """(B, steps) int64 ids (any device) -> list of B strings, words up to (excluding) the first <end>
    (`Decoder.decode_tokens`, layer.py:464-477), with ONE device->host copy for the whole table."""
    ids = token_table.detach().to('cpu', non_blocking=False).tolist()
    idx2word = decoder.vocab.idx2word
    end = decoder.vocab('<end>')
    out = []
    for row in ids:
        words = []
        for t in row:
            if t == end:
                break
            words.append(idx2word[t])
        out.append(' '.join(words))
    return out


def gather_results(net, opt, eval_loader, multi_modal=False, multi_gpu=False, device=None):
    """Same signature and return value as the reference's `evaluate.gather_results` (evaluate.py:101-116): an OrderedDict
    video id -> caption for the batches `eval_loader` yields on THIS rank, and the list of attention tables."""
    core = net.module if multi_gpu and hasattr(net, 'module') else net
    if device is None:
        device = next(core.parameters()).device
    result = collections.OrderedDict()
    alpha_all = []
    with torch.no_grad():
        for frames, regions, _spatials, video_ids in eval_loader:
            frames = frames.to(device, non_blocking=True)
            regions = regions[:, :, :opt.num_obj, :].to(device, non_blocking=True)
            outputs, _, _, alpha = net(frames, regions, None)
            alpha_all.append(alpha)
            for vid, s in zip(video_ids, tokens_to_captions(core.decoder, outputs)):
                result[int(vid) if torch.is_tensor(vid) else vid] = s
    return result, alpha_all


def merge_rank_results(local_result, process_group=None):
    """All ranks' caption dicts merged on every rank (rank order, later ranks win on duplicates - the reference's merge at
    run_gun.py:270-276, for any number of ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return collections.OrderedDict(local_result)
    world = dist.get_world_size(process_group)
    parts = [None] * world
    dist.all_gather_object(parts, dict(local_result), group=process_group)
    merged = collections.OrderedDict()
    for p in parts:
        merged.update(p)
    return merged


def gather_results_all_ranks(net, opt, eval_loader, multi_gpu=False, process_group=None):
    """Decode this rank's shard, then merge: (merged captions of the whole evaluation set, this rank's attention tables)."""
    local, alpha_all = gather_results(net, opt, eval_loader, multi_modal=True, multi_gpu=multi_gpu)
    return merge_rank_results(local, process_group), alpha_all
