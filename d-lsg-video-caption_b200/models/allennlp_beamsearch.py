"""Drop-in mirror of the reference's models/allennlp_beamsearch.py (BeamSearch.__init__ :39-49, .search :51-294).

Same constructor, same `search(start_predictions, start_state, step)` contract and return shapes
((batch, beam, steps_taken) int64, (batch, beam) fp32).  The per-step candidate selection runs in the
libdlsg beam kernels (warp-level top-k with the after-<end> forcing fused in, candidate merge with
back-pointers, back-track); arbitrary user state dicts are re-indexed by back-pointer with an index
gather.  Decoder.forward does not go through this generic entry: it uses dlsg.decoder.decode_beam, which
keeps the LSTM state in the kernels' operand buffers.
"""
from typing import List, Callable, Tuple, Dict
import warnings

import torch

from dlsg.errors import ConfigurationError
from dlsg import ops

StateType = Dict[str, torch.Tensor]
StepFunctionType = Callable[[torch.Tensor, StateType], Tuple[torch.Tensor, StateType]]


class BeamSearch:
    def __init__(self, end_index: int, max_steps: int = 50, beam_size: int = 10, per_node_beam_size: int = None) -> None:
        self._end_index = end_index
        self.max_steps = max_steps
        self.beam_size = beam_size
        self.per_node_beam_size = per_node_beam_size or beam_size

    def search(self, start_predictions: torch.Tensor, start_state: StateType, step: StepFunctionType
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        be = ops.backend()
        B = start_predictions.size()[0]
        beam, k, end = self.beam_size, self.per_node_beam_size, self._end_index
        dev = start_predictions.device
        logp0, state = step(start_predictions, start_state)
        V = logp0.size()[1]
        if k > V:
            raise ConfigurationError(
                f"Target vocab size ({V:d}) too small relative to per_node_beam_size ({k:d}).\n"
                f"Please decrease beam_size or per_node_beam_size.")
        T = self.max_steps
        preds = torch.empty((T, B, beam), dtype=torch.int64, device=dev)
        backs = torch.empty((max(T - 1, 1), B, beam), dtype=torch.int64, device=dev)
        lps = torch.empty((2, B, beam), dtype=torch.float32, device=dev)
        be.beam_topk(logp0.float().contiguous(), None, end, beam, lps[0], preds[0], normalize=False)
        if beam == 1 and bool((preds[0] == end).all()):
            warnings.warn("Empty sequences predicted. You may want to increase the beam size or ensure "
                          "your step function is working properly.", RuntimeWarning)
            return preds[0].unsqueeze(-1), lps[0]
        for key, st in state.items():
            _, *last = st.size()
            state[key] = st.unsqueeze(1).expand(B, beam, *last).reshape(B * beam, *last)
        top_lp = torch.empty((B * beam, k), dtype=torch.float32, device=dev)
        top_id = torch.empty((B * beam, k), dtype=torch.int64, device=dev)
        base = (torch.arange(B, device=dev) * beam).unsqueeze(1)
        cur, S = 0, T
        for s in range(1, T):
            last_predictions = preds[s - 1].reshape(B * beam)
            if bool((last_predictions == end).all()):
                S = s
                break
            logp, state = step(last_predictions, state)
            be.beam_topk(logp.float().contiguous(), last_predictions, end, k, top_lp, top_id, normalize=False)
            be.beam_merge(top_lp, top_id, lps[cur], B, beam, k, lps[1 - cur], preds[s], backs[s - 1], None, end)
            cur = 1 - cur
            rows = (base + backs[s - 1]).reshape(-1)
            for key, st in state.items():
                state[key] = st.index_select(0, rows)
        last_lp = lps[cur]
        if not bool(torch.isfinite(last_lp).all()):
            warnings.warn("Infinite log probabilities encountered. Some final sequences may not make sense. "
                          "This can happen when the beam size is larger than the number of valid (non-zero "
                          "probability) transitions that the step function produces.", RuntimeWarning)
        out = torch.empty((B, beam, S), dtype=torch.int64, device=dev)
        if S == 1:
            out.copy_(preds[0].unsqueeze(-1))
        else:
            be.beam_backtrack(preds, backs, S, B, beam, out)
        return out, last_lp
