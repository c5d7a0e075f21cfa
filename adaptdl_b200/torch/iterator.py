"""Elastic BPTT iterator for language modelling on one long token stream.

Same role and constructor shape as the reference's
``AdaptiveBPTTIterator`` (``torch/iterator.py:33-121``), without the
legacy-torchtext dependency: the dataset is a 1-D tensor (or sequence) of
token ids, or any object with a ``tokens`` attribute / torchtext-style
``dataset[0].text`` + ``fields['text'].numericalize``.

The stream is folded into ``[len/bsz, bsz]`` columns for the *current* local
batch size; replica r reads rows ``start + r*bptt_len`` stepping by
``bptt_len * replicas``. After a rescale the local batch size (hence the
fold) changes, so the resume row is re-derived proportionally. All replicas
iterate the number of steps of the highest rank so per-step collectives stay
symmetric.
"""

import collections
import math

import torch

from adaptdl_b200 import env
from adaptdl_b200.torch.data import AdaptiveDataLoaderMixin

__all__ = ["AdaptiveBPTTIterator", "Batch"]

Batch = collections.namedtuple("Batch", ["text", "target"])


def _token_stream(dataset):
    if torch.is_tensor(dataset):
        return dataset.reshape(-1).long()
    tokens = getattr(dataset, "tokens", None)
    if tokens is not None:
        return torch.as_tensor(tokens).reshape(-1).long()
    try:                                   # torchtext-legacy style
        text = dataset[0].text
        field = dataset.fields["text"]
        return field.numericalize([text]).reshape(-1).long()
    except (AttributeError, KeyError, TypeError, IndexError):
        return torch.as_tensor(list(dataset)).reshape(-1).long()


class AdaptiveBPTTIterator(AdaptiveDataLoaderMixin):
    """Arguments:
        dataset: token stream (see module docstring).
        batch_size (int): target *global* batch size (number of columns).
        bptt_len (int): tokens per back-propagation-through-time window.
        max_batch_size, local_bsz_bounds: enable adaptive batch size.
        device: where batches are placed.
        batch_first (bool): yield ``[bsz, seq]`` instead of ``[seq, bsz]``.
        repeat (bool): iterate forever.
        pad_token (int): id used to pad the stream to a whole fold.
    """

    def __init__(self, dataset, batch_size, bptt_len, **kwargs):
        max_batch_size = kwargs.pop("max_batch_size", None)
        local_bsz_bounds = kwargs.pop("local_bsz_bounds", None)
        self.device = kwargs.pop("device", None)
        self.batch_first = kwargs.pop("batch_first", False)
        self.repeat = kwargs.pop("repeat", False)
        self.pad_token = kwargs.pop("pad_token", 0)
        kwargs.pop("train", None), kwargs.pop("shuffle", None)
        kwargs.pop("sort", None)
        if kwargs:
            raise TypeError("unexpected arguments: {}".format(sorted(kwargs)))
        self.dataset = dataset
        self.batch_size = batch_size
        self.bptt_len = bptt_len
        self.iterations = 0
        AdaptiveDataLoaderMixin.__init__(self, batch_size)
        self.num_replicas = env.num_replicas()
        self.rank = env.replica_rank()
        self._tokens = None
        if max_batch_size and local_bsz_bounds:
            self._elastic.autoscale_batch_size(max_batch_size,
                                               local_bsz_bounds)

    @staticmethod
    def _recompute_start(prev_curr, prev_end, curr_end):
        """Resume row after the fold width changed."""
        if prev_end == 0:
            return prev_curr
        return math.ceil(prev_curr * curr_end / prev_end)

    def _fold(self, local_bsz):
        if self._tokens is None:
            self._tokens = _token_stream(self.dataset)
        tokens = self._tokens
        rows = math.ceil(tokens.numel() / local_bsz)
        pad = rows * local_bsz - tokens.numel()
        if pad:
            tokens = torch.cat([tokens, tokens.new_full((pad,),
                                                        self.pad_token)])
        data = tokens.view(local_bsz, -1).t().contiguous()
        if self.device is not None:
            data = data.to(self.device)
        return data

    def __len__(self):
        local = self._elastic.current_local_bsz or math.ceil(
            self.batch_size / self.num_replicas)
        rows = math.ceil(_token_stream(self.dataset).numel() / local)
        return math.ceil((rows - 1) / (self.bptt_len * self.num_replicas))

    def __iter__(self):
        elastic = self._elastic
        with elastic.context():
            if elastic.skipdone():
                return
            self.batch_size = elastic._sync_local_bsz()
            data = self._fold(self.batch_size)
            end = data.size(0)
            elastic.current_index = self._recompute_start(
                elastic.current_index, elastic.end_index, end)
            elastic.end_index = end
            step = self.bptt_len * self.num_replicas
            self.iterations = 0
            while True:
                first = elastic.current_index
                start = first + self.bptt_len * self.rank
                highest = first + self.bptt_len * (self.num_replicas - 1)
                # the last window needs at least one target token
                steps = max(math.ceil((end - 1 - highest) / step), 0)
                for n, i in enumerate(range(start, end, step)):
                    if n >= steps:
                        break
                    self.iterations += 1
                    with elastic.profile(self.training and i > 0):
                        seq_len = min(self.bptt_len, end - i - 1)
                        assert seq_len > 0
                        text = data[i:i + seq_len]
                        target = data[i + 1:i + 1 + seq_len]
                        if self.batch_first:
                            text = text.t().contiguous()
                            target = target.t().contiguous()
                        yield Batch(text, target)
                        elastic.current_index += step
                if not self.repeat:
                    break
                elastic.current_index = 0
