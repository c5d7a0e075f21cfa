"""``torchtext.datasets.WikiText2.splits(TEXT)`` of legacy torchtext, without
the download: three one-example language-modelling datasets over a small
synthetic vocabulary (sizes from ``ADL_TEST_FAKE_TOKENS``, default 4000 train
tokens)."""
import os
import random

from . import data


class WikiText2(data.Dataset):
    @classmethod
    def splits(cls, text_field, **kwargs):
        rng = random.Random(0)
        words = ["w%d" % i for i in range(200)]
        n = int(os.environ.get("ADL_TEST_FAKE_TOKENS", "4000"))
        out = []
        for size in (n, max(n // 8, 200), max(n // 8, 200)):
            # Zipf-like draw so the model has something to learn
            tokens = [words[min(int(rng.paretovariate(1.2)) - 1, 199)]
                      for _ in range(size)]
            fields = [("text", text_field)]
            out.append(cls([data.Example.fromlist([tokens], fields)],
                           fields))
        return tuple(out)
