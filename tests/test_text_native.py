"""csrc/text_native.cpp (the analysis chains in native code) against kaito_b200/text.py (the Python restatement, itself
checked against transformers' BertTokenizer and Snowball vectors in test_wordpiece.py / test_text.py): identical output on
fuzzed ASCII text -- every Porter2 rule family, stop words, digits/underscores, punctuation, control characters."""
import numpy as np
import pytest

from kaito_b200 import text as T

PREFIX = ["gener", "commun", "arsen", "hap", "rel", "nation", "oper", "sky", "cri", "tr", "st", "bl", "fl", "agr", "cond", "y", "by", "say", "succ", "proc",
          "exc", "inn", "out", "cann", "herr", "earr", "ug", "on", "ear", "gent", "id", "sing", "new", "how", "atl", "cosm", "bi", "and", "d", "l", "t"]
MIDDLE = ["", "a", "e", "i", "o", "u", "y", "at", "iz", "bl", "ional", "ic", "al", "ous", "iv", "ful", "less", "ent", "abl", "og", "l", "ll", "ss", "tt", "pp", "ee"]
SUFFIX = ["", "s", "es", "ies", "ied", "sses", "us", "ss", "ed", "ing", "edly", "ingly", "eed", "eedly", "y", "ly", "li", "ization", "ational", "fulness",
          "ousness", "iveness", "tional", "biliti", "lessli", "entli", "ation", "alism", "aliti", "ousli", "iviti", "fulli", "enci", "anci", "abli",
          "izer", "ator", "alli", "bli", "ogi", "alize", "icate", "iciti", "ative", "ical", "ness", "ful", "ement", "ance", "ence", "able", "ible",
          "ment", "ant", "ent", "ism", "ate", "iti", "ous", "ive", "ize", "ion", "sion", "tion", "al", "er", "ic", "e", "le", "ll", "l", "'s", "'", "'s'"]
EXACT = ["skis", "skies", "dying", "lying", "tying", "idly", "gently", "ugly", "early", "only", "singly", "sky", "news", "howe", "atlas", "cosmos",
         "bias", "andes", "inning", "outing", "canning", "herring", "earring", "proceed", "exceed", "succeed", "the", "and", "will", "THE", "Into",
         "a", "I", "x1", "a_b", "__init__", "v2_final", "2024", "3.14", "e-mail", "don't", "it's", "O'Reilly", "KAITO", "GPUs", "x"]
PUNCT = list(" \t\n\r.,;:!?()[]{}<>\"'`~@#$%^&*-+=/\\|_") + ["\x00", "\x01", "\x0b", "\x0c", "\x1f", "\x7f"]


def _texts(seed, n, words_per_text):
    g = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        parts = []
        for _ in range(int(g.integers(0, words_per_text))):
            r = g.random()
            if r < 0.2:
                w = EXACT[int(g.integers(len(EXACT)))]
            else:
                w = PREFIX[int(g.integers(len(PREFIX)))] + MIDDLE[int(g.integers(len(MIDDLE)))] + SUFFIX[int(g.integers(len(SUFFIX)))]
                if g.random() < 0.15:
                    w = w.upper() if g.random() < 0.5 else w.capitalize()
            parts.append(w)
            parts.append("".join(PUNCT[int(g.integers(len(PUNCT)))] for _ in range(int(g.integers(1, 3)))))
        out.append("".join(parts))
    return out


needs_native = pytest.mark.skipif(not T._native_lib(), reason="libkaito_rag.so not built")


@needs_native
def test_native_analysis_equals_python_spec():
    texts = _texts(1, 1500, 40) + ["", " ", "a", "ab", "the", "__", "x" * 300, "'s", "ies", "ied", "sses"]
    n_tok = 0
    for t in texts:
        a, b = T.tokenize(t), T.tokenize_py(t)
        assert a == b, (t, a, b)
        n_tok += len(b)
    assert n_tok > 15000
    # every generated word on its own (no context effects) -- exercises each suffix rule with each stem shape
    for p in PREFIX:
        for m in MIDDLE:
            for s in SUFFIX:
                w = p + m + s
                assert T.tokenize(w) == T.tokenize_py(w), w


@needs_native
def test_non_ascii_text_takes_the_python_path():
    t = "Kubernetes opérateurs naïve 北京 running"
    assert T.tokenize(t) == T.tokenize_py(t) and "run" in T.tokenize(t)


def _vocab(seed):
    g = np.random.default_rng(seed)
    toks = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + list("abcdefghijklmnopqrstuvwxyz0123456789") + ["##" + c for c in "abcdefghijklmnopqrstuvwxyz0123456789"]
    toks += list(".,;:!?()[]{}<>\"'`~@#$%^&*-+=/\\|_")
    pieces = set()
    for w in [p + m + s for p in PREFIX for m in MIDDLE[:8] for s in SUFFIX[:12]]:
        w = w.replace("'", "").lower()
        if len(w) >= 2 and g.random() < 0.3:
            cut = int(g.integers(1, len(w)))
            pieces.add(w[:cut]); pieces.add("##" + w[cut:])
        if g.random() < 0.1:
            pieces.add(w)
    return toks + sorted(pieces - set(toks))


@needs_native
@pytest.mark.parametrize("lower", [True, False])
def test_native_wordpiece_equals_python_spec(lower):
    tok = T.WordPieceTokenizer(_vocab(3), do_lower_case=lower, max_length=64)
    assert tok._nat is not None
    texts = _texts(5, 800, 30) + ["", "   ", "x" * 150, "a" * 99 + "!", "unknownzzzzqqq", "[CLS] literal", "\x00\x01ab\x7fcd"]
    got = tok.encode_batch(texts)
    for t, ids in zip(texts, got):
        assert ids == tok.encode_py(t), (t, ids, tok.encode_py(t))
        assert ids[0] == tok.cls and ids[-1] == tok.sep and len(ids) <= 64
    assert any(len(i) == 64 for i in got) and any(tok.unk in i for i in got)
    mixed = ["plain ascii", "naïve café", "operators running"]             # the non-ASCII text goes through encode_py
    assert tok.encode_batch(mixed) == [tok.encode_py(t) for t in mixed]
    # a duplicated vocabulary entry keeps its LAST id, as the Python dict does
    dup = T.WordPieceTokenizer(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "ab", "cd", "ab"], max_length=8)
    assert dup.encode("ab cd") == dup.encode_py("ab cd") == [2, 6, 5, 3]


@needs_native
def test_native_analysis_is_faster_than_python():
    import time
    texts = _texts(9, 300, 200)
    t0 = time.perf_counter(); a = [T.tokenize(t) for t in texts]; t1 = time.perf_counter()
    b = [T.tokenize_py(t) for t in texts]; t2 = time.perf_counter()
    assert a == b and (t1 - t0) * 3 < (t2 - t1), ((t1 - t0), (t2 - t1))
