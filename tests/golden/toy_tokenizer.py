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
