"""A deterministic whitespace tokenizer with the three methods the p2p helpers use (encode / decode /
model_max_length), standing in for CLIPTokenizer (no tokenizer files are available offline)."""


class ToyTokenizer:
    model_max_length = 77
    BOS, EOS = 1, 2

    def __init__(self):
        self.vocab = {}
        self.inv = {self.BOS: "<s>", self.EOS: "</s>"}

    def _id(self, piece):
        if piece not in self.vocab:
            self.vocab[piece] = len(self.vocab) + 3
            self.inv[self.vocab[piece]] = piece
        return self.vocab[piece]

    def encode(self, text):
        ids = [self.BOS]
        for word in text.split(" "):
            if not word:
                continue
            # words longer than 6 characters are split into two pieces (exercises multi-token words)
            pieces = [word] if len(word) <= 6 else [word[:4], word[4:]]
            ids += [self._id(p) for p in pieces]
        return ids + [self.EOS]

    def decode(self, ids):
        return "".join(self.inv[i] for i in ids)


class ToyCallableTokenizer(ToyTokenizer):
    """Adds the `tokenizer(captions, padding=..., max_length=..., truncation=..., return_tensors="pt").input_ids`
    call the SDXL prompt encoder uses."""

    def __call__(self, captions, padding="max_length", max_length=77, truncation=True, return_tensors="pt"):
        import torch
        from types import SimpleNamespace
        rows = []
        for c in captions:
            ids = self.encode(c)[:max_length]
            rows.append(ids + [self.EOS] * (max_length - len(ids)))
        return SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.long))


class ToyTextEncoder:
    """Deterministic stand-in for CLIPTextModel(WithProjection): `enc(ids, output_hidden_states=True)` returns an
    object with `[0]` (pooled / last hidden state) and `.hidden_states` (list of 3 tensors)."""

    def __init__(self, width, seed, pooled_dim=None):
        import torch
        g = torch.Generator().manual_seed(seed)
        self.table = torch.randn(4096, width, generator=g)
        self.pool = torch.randn(width, pooled_dim, generator=g) if pooled_dim else None
        self.device = torch.device("cpu")

    def __call__(self, ids, output_hidden_states=True):
        import torch
        h0 = self.table[ids]
        h1 = torch.tanh(h0 + h0.mean(1, keepdim=True))
        h2 = h1 * 0.5 + 0.1
        first = h2[:, 0] @ self.pool if self.pool is not None else h2

        class Out(tuple):
            pass
        out = Out((first,))
        out.hidden_states = [h0, h1, h2]
        return out
