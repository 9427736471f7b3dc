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
