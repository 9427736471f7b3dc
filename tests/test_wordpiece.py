"""WordPiece restatement against transformers.BertTokenizer (slow, pure Python) built from the same
synthetic vocab.txt -- the real bge vocab is not available offline."""
import os

import pytest

from kaito_b200.text import WordPieceTokenizer

VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "the", "quick", "brown", "fox", "jump", "##s", "##ed", "##ing", "over", "lazy",
         "dog", ",", ".", "!", "?", "'", "-", "un", "##believ", "##able", "retrie", "##val", "engine", "gpu", "b", "##200", "200", "a",
         "cafe", "naive", "##ly", "中", "文", "token", "##izer", "##ization", "hello", "world", "##s", "1", "##2", "##3", "(", ")", "re", "##present",
         "this", "question", "for", "search", "relevant", "passage", ":", "##e", "x"]

TEXTS = ["The quick brown fox jumps over the lazy dog.", "Unbelievable retrieval engine, GPU B200!", "café naïvely tokenizers?",
         "hello   world\t123 (unknownword) x-x", "中文 token", "", "Represent this question for searching relevant passages: what's b200"]


def test_wordpiece_matches_transformers(tmp_path):
    transformers = pytest.importorskip("transformers")
    p = tmp_path / "vocab.txt"
    seen, vocab = set(), []
    for t in VOCAB:
        if t not in seen:
            seen.add(t); vocab.append(t)
    p.write_text("\n".join(vocab) + "\n", encoding="utf-8")
    ref = transformers.BertTokenizer(str(p), do_lower_case=True)
    mine = WordPieceTokenizer.from_file(str(p))
    for t in TEXTS:
        assert mine.encode(t) == ref.encode(t, add_special_tokens=True), t
    long = "fox " * 600
    assert mine.encode(long) == ref.encode(long, add_special_tokens=True, truncation=True, max_length=512)
    assert len(mine.encode(long)) == 512


def test_wordpiece_fuzz_against_the_fast_tokenizer(tmp_path):
    """the reference's tokenizer is the Rust one (sentence-transformers -> AutoTokenizer -> BertTokenizerFast,
    embedding/huggingface_local_embedding.py:34-53): random text over an alphabet with accents, CJK, punctuation, control and
    zero-width characters, digits and over-long words must give the same ids (native and Python paths of kaito_b200.text)"""
    transformers = pytest.importorskip("transformers")
    pytest.importorskip("tokenizers")
    import random
    seen, vocab = set(), []
    letters = "abcdefghijklmnopqrstuvwxyz"
    for t in VOCAB + list(letters) + ["##" + c for c in letters] + [str(d) for d in range(10)] + ["##" + str(d) for d in range(10)] + \
            ["##ab", "##cd", "ab", "cd", "ing", "##ion", "tion", "日", "本", "語", "é", "##é", "ü", "$", "%", "&", "/", ";", "@", "[", "]", "_", "~"]:
        if t not in seen:
            seen.add(t); vocab.append(t)
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(vocab) + "\n", encoding="utf-8")
    fast = transformers.BertTokenizer(str(p), do_lower_case=True)       # transformers >= 5: backed by the Rust `tokenizers` WordPiece
    assert fast.convert_ids_to_tokens(fast.encode("a]b")) == ["[CLS]", "a", "]", "b", "[SEP]"]     # the vocabulary really loaded
    mine = WordPieceTokenizer.from_file(str(p))
    rnd = random.Random(1234)
    alphabet = list(letters) + list("ABCXYZ") + list("0123456789") + list(" \t\n  ") + list(".,!?'-()[]$%&/;:@_~") + \
        list("éèüñçÅøß") + list("中文日本語") + ["​", "­", "\x00", "\x07", "�", "é", "　", "İ", "ǅ"]
    texts = []
    for _ in range(400):
        n = rnd.randint(0, 60)
        texts.append("".join(rnd.choice(alphabet) for _ in range(n)))
    texts += ["a" * 101, "a" * 100, "ab" * 60 + " cd", " ".join(["ing"] * 700), "x" * 99 + "é"]
    for t in texts:
        want = fast.encode(t, add_special_tokens=True, truncation=True, max_length=512)
        assert mine.encode(t) == want, repr(t)
        assert mine.encode_py(t) == want, repr(t)
    flat, offs = mine.encode_batch_flat(texts)
    assert [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(texts))] == [fast.encode(t, add_special_tokens=True, truncation=True, max_length=512) for t in texts]
