import collections

import torch

from . import utils  # noqa: F401


class Field(object):
    def __init__(self, tokenize=None, init_token=None, eos_token=None,
                 pad_token="<pad>", unk_token="<unk>", sequential=True,
                 lower=False, batch_first=False):
        self.tokenize = tokenize or str.split
        self.init_token, self.eos_token = init_token, eos_token
        self.pad_token, self.unk_token = pad_token, unk_token
        self.sequential, self.lower = sequential, lower
        self.batch_first = batch_first
        self.vocab = None

    def preprocess(self, value):
        if self.sequential and isinstance(value, str):
            value = self.tokenize(value.rstrip("\n"))
        if self.lower:
            value = [tok.lower() for tok in value]
        return value

    def build_vocab(self, *datasets):
        counter = collections.Counter()
        for dataset in datasets:
            for example in dataset:
                for name, field in dataset.fields.items():
                    if field is self:
                        counter.update(getattr(example, name))
        specials = [tok for tok in (self.unk_token, self.pad_token,
                                    self.init_token, self.eos_token)
                    if tok is not None]
        itos = specials + [tok for tok, _ in counter.most_common()
                           if tok not in specials]
        self.vocab = collections.namedtuple("Vocab", ["itos", "stoi"])(
            itos, {tok: i for i, tok in enumerate(itos)})

    def numericalize(self, batch, device=None):
        unk = self.vocab.stoi.get(self.unk_token, 0)
        ids = [[self.vocab.stoi.get(tok, unk) for tok in row]
               for row in batch]
        out = torch.tensor(ids, dtype=torch.long, device=device)
        return out if self.batch_first else out.t().contiguous()


class Example(object):
    @classmethod
    def fromlist(cls, data, fields):
        example = cls()
        for value, (name, field) in zip(data, fields):
            if field is not None:
                setattr(example, name, field.preprocess(value))
        return example


class Dataset(object):
    def __init__(self, examples, fields):
        self.examples = list(examples)
        self.fields = dict(fields)

    def __getitem__(self, i):
        return self.examples[i]

    def __len__(self):
        return len(self.examples)

    def __iter__(self):
        return iter(self.examples)
