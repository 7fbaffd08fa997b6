"""Token alignment between an original and an edited prompt (host-side, O(77^2), once per edit).

Behaviour follows utils/seq_aligner.py of the reference (global Needleman-Wunsch alignment with gap 0, match +1,
mismatch -1 and the tie-break order left > up > diagonal, utils/seq_aligner.py:50-67), because the edit
controllers (p2p.AttentionRefine / AttentionReplace) consume its mappers; it is not on the accelerated path.
"""
import numpy as np
import torch

GAP, MATCH, MISMATCH = 0, 1, -1
_LEFT, _UP, _DIAG, _STOP = 1, 2, 3, 4


def _align_tables(a, b):
    """DP score table and trace-back moves for aligning token lists a (rows) and b (columns)."""
    na, nb = len(a), len(b)
    score = np.zeros((na + 1, nb + 1), dtype=np.int32)
    move = np.zeros((na + 1, nb + 1), dtype=np.int32)
    score[0, :] = np.arange(nb + 1) * GAP
    score[:, 0] = np.arange(na + 1) * GAP
    move[0, 1:], move[1:, 0], move[0, 0] = _LEFT, _UP, _STOP
    for i in range(1, na + 1):
        for j in range(1, nb + 1):
            cand = (score[i, j - 1] + GAP, score[i - 1, j] + GAP,
                    score[i - 1, j - 1] + (MATCH if a[i - 1] == b[j - 1] else MISMATCH))
            best = max(cand)
            score[i, j] = best
            move[i, j] = _LEFT if best == cand[0] else (_UP if best == cand[1] else _DIAG)
    return score, move


def _walk_back(a, b, move):
    """Pairs (index in b, index in a or -1) in increasing order of the b index."""
    i, j, pairs = len(a), len(b), []
    while i > 0 or j > 0:
        m = move[i, j]
        if m == _DIAG:
            i, j = i - 1, j - 1
            pairs.append((j, i))
        elif m == _LEFT:
            j -= 1
            pairs.append((j, -1))
        elif m == _UP:
            i -= 1
        else:
            break
    return torch.tensor(pairs[::-1], dtype=torch.int64)


def get_mapper(x: str, y: str, tokenizer, max_len=77):
    ids_x, ids_y = tokenizer.encode(x), tokenizer.encode(y)
    _, move = _align_tables(ids_x, ids_y)
    base = _walk_back(ids_x, ids_y, move)
    n = base.shape[0]
    alphas = torch.ones(max_len)
    alphas[:n] = base[:, 1].ne(-1).float()
    mapper = torch.zeros(max_len, dtype=torch.int64)
    mapper[:n] = base[:, 1]
    mapper[n:] = len(ids_y) + torch.arange(max_len - len(ids_y))
    return mapper, alphas


def get_refinement_mapper(prompts, tokenizer, max_len=77):
    pairs = [get_mapper(prompts[0], p, tokenizer, max_len) for p in prompts[1:]]
    return torch.stack([m for m, _ in pairs]), torch.stack([a for _, a in pairs])


def get_word_inds(text: str, word_place, tokenizer):
    """Token positions (1-based, after BOS) of the word(s) `word_place` (index or literal word) in `text`."""
    words = text.split(" ")
    if isinstance(word_place, str):
        wanted = [i for i, w in enumerate(words) if w == word_place]
    elif isinstance(word_place, int):
        wanted = [word_place]
    else:
        wanted = list(word_place)
    found = []
    if wanted:
        pieces = [tokenizer.decode([t]).strip("#") for t in tokenizer.encode(text)][1:-1]
        consumed, w = 0, 0
        for pos, piece in enumerate(pieces):
            consumed += len(piece)
            if w in wanted:
                found.append(pos + 1)
            if consumed >= len(words[w]):
                w, consumed = w + 1, 0
    return np.array(found)


def _replacement_mapper(x: str, y: str, tokenizer, max_len=77):
    wx, wy = x.split(" "), y.split(" ")
    if len(wx) != len(wy):
        raise ValueError(f"attention replacement edit can only be applied on prompts with the same length"
                         f" but prompt A has {len(wx)} words and prompt B has {len(wy)} words.")
    changed = [k for k in range(len(wy)) if wy[k] != wx[k]]
    src = [get_word_inds(x, k, tokenizer) for k in changed]
    dst = [get_word_inds(y, k, tokenizer) for k in changed]
    m = np.zeros((max_len, max_len))
    i = j = nxt = 0
    while i < max_len and j < max_len:
        if nxt < len(src) and src[nxt][0] == i:
            s, d = src[nxt], dst[nxt]
            if len(s) == len(d):
                m[s, d] = 1
            else:
                for col in d:
                    m[s, col] = 1 / len(d)
            nxt += 1
            i, j = i + len(s), j + len(d)
        elif nxt < len(src):
            m[i, j] = 1
            i, j = i + 1, j + 1
        else:
            m[j, j] = 1
            i, j = i + 1, j + 1
    return torch.from_numpy(m).float()


def get_replacement_mapper(prompts, tokenizer, max_len=77):
    return torch.stack([_replacement_mapper(prompts[0], p, tokenizer, max_len) for p in prompts[1:]])
