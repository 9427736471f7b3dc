"""Host-side text pipeline for the sparse side of /retrieve.

The reference tokenises inside `BM25Retriever.from_defaults` (hybrid_retriever.py:122-125) with
bm25s + PyStemmer, neither vendored nor installable offline [3P-unverified]; this module
restates that pipeline from the published algorithms:

  bm25s.tokenize(texts, stopwords="english", stemmer=Stemmer.Stemmer("english")):
    lower-case, regex r"(?u)\\b\\w\\w+\\b", drop the 33-word English stop list, then stem
    each remaining token with the Snowball English ("Porter2") stemmer.

Term ids are assigned by `Vocabulary` in first-seen order; they are what the C ABI takes
(`term_ids`, `q_terms` in include/kaito_rag.h).  Text stays on the host.
"""
from __future__ import annotations

import re
from collections import Counter

import numpy as np

TOKEN_PATTERN = re.compile(r"(?u)\b\w\w+\b")

# bm25s.stopwords.STOPWORDS_EN
STOPWORDS_EN = frozenset((
    "a", "an", "and", "are", "as", "at", "be", "but", "by", "for", "if", "in", "into", "is", "it", "no", "not", "of",
    "on", "or", "such", "that", "the", "their", "then", "there", "these", "they", "this", "to", "was", "will", "with",
))

# ---------------------------------------------------------------- Snowball English stemmer
_VOWELS = "aeiouy"
_DOUBLES = ("bb", "dd", "ff", "gg", "mm", "nn", "pp", "rr", "tt")
_LI_ENDING = "cdeghkmnrt"
_EXCEPTIONS = {"skis": "ski", "skies": "sky", "dying": "die", "lying": "lie", "tying": "tie", "idly": "idl",
               "gently": "gentl", "ugly": "ugli", "early": "earli", "only": "onli", "singly": "singl", "sky": "sky",
               "news": "news", "howe": "howe", "atlas": "atlas", "cosmos": "cosmos", "bias": "bias", "andes": "andes"}
_EXCEPTIONS_1A = frozenset(("inning", "outing", "canning", "herring", "earring", "proceed", "exceed", "succeed"))
_STEP2 = (("ization", "ize"), ("ational", "ate"), ("fulness", "ful"), ("ousness", "ous"), ("iveness", "ive"),
          ("tional", "tion"), ("biliti", "ble"), ("lessli", "less"), ("entli", "ent"), ("ation", "ate"),
          ("alism", "al"), ("aliti", "al"), ("ousli", "ous"), ("iviti", "ive"), ("fulli", "ful"), ("enci", "ence"),
          ("anci", "ance"), ("abli", "able"), ("izer", "ize"), ("ator", "ate"), ("alli", "al"), ("bli", "ble"),
          ("ogi", "og"), ("li", ""))
_STEP3 = (("ational", "ate"), ("tional", "tion"), ("alize", "al"), ("icate", "ic"), ("iciti", "ic"), ("ative", ""),
          ("ical", "ic"), ("ness", ""), ("ful", ""))
_STEP4 = ("ement", "ance", "ence", "able", "ible", "ment", "ant", "ent", "ism", "ate", "iti", "ous", "ive", "ize",
          "ion", "al", "er", "ic")


def _is_vowel(w: str, i: int) -> bool:
    return w[i] in _VOWELS


def _region_after_vc(w: str, start: int) -> int:
    """index after the first non-vowel that follows a vowel, searching from `start`"""
    for i in range(start + 1, len(w)):
        if not _is_vowel(w, i) and _is_vowel(w, i - 1):
            return i + 1
    return len(w)


def _r1_r2(w: str):
    if w.startswith(("gener", "arsen")):
        r1 = 5
    elif w.startswith("commun"):
        r1 = 6
    else:
        r1 = _region_after_vc(w, 0)
    return r1, _region_after_vc(w, r1)


def _ends_short_syllable(w: str) -> bool:
    n = len(w)
    if n == 2:
        return _is_vowel(w, 0) and not _is_vowel(w, 1)
    if n >= 3:
        return (not _is_vowel(w, n - 3)) and _is_vowel(w, n - 2) and (not _is_vowel(w, n - 1)) and w[-1] not in "wxY"
    return False


def _contains_vowel(w: str) -> bool:
    return any(c in _VOWELS for c in w)


def stem(word: str) -> str:
    """Snowball English (Porter2) stem of a lower-case token."""
    if len(word) <= 2:
        return word
    if word in _EXCEPTIONS:
        return _EXCEPTIONS[word]
    w = word[1:] if word.startswith("'") else word
    # mark consonant y
    chars = list(w)
    for i, c in enumerate(chars):
        if c == "y" and (i == 0 or chars[i - 1] in _VOWELS):
            chars[i] = "Y"
    w = "".join(chars)
    r1, r2 = _r1_r2(w)

    # step 0
    for suf in ("'s'", "'s", "'"):
        if w.endswith(suf):
            w = w[: -len(suf)]
            break
    # step 1a
    if w.endswith("sses"):
        w = w[:-2]
    elif w.endswith(("ied", "ies")):
        w = w[:-2] if len(w) > 4 else w[:-1]
    elif w.endswith(("us", "ss")):
        pass
    elif w.endswith("s"):
        if _contains_vowel(w[:-2]):
            w = w[:-1]
    if w in _EXCEPTIONS_1A:
        return w.replace("Y", "y")
    # step 1b
    if w.endswith("eedly"):
        if len(w) - 5 >= r1:
            w = w[:-3]
    elif w.endswith("eed"):
        if len(w) - 3 >= r1:
            w = w[:-1]
    else:
        for suf in ("ingly", "edly", "ing", "ed"):
            if w.endswith(suf):
                stem_part = w[: -len(suf)]
                if _contains_vowel(stem_part):
                    w = stem_part
                    if w.endswith(("at", "bl", "iz")):
                        w += "e"
                    elif w.endswith(_DOUBLES):
                        w = w[:-1]
                    elif _ends_short_syllable(w) and r1 >= len(w):
                        w += "e"
                break
    # step 1c
    if len(w) > 2 and w[-1] in "yY" and w[-2] not in _VOWELS:
        w = w[:-1] + "i"
    # step 2
    for suf, rep in _STEP2:
        if w.endswith(suf):
            if len(w) - len(suf) >= r1:
                if suf == "ogi":
                    if w[: -len(suf)].endswith("l"):
                        w = w[: -len(suf)] + rep
                elif suf == "li":
                    if len(w) > 2 and w[-3] in _LI_ENDING:
                        w = w[:-2]
                else:
                    w = w[: -len(suf)] + rep
            break
    # step 3
    for suf, rep in _STEP3:
        if w.endswith(suf):
            if len(w) - len(suf) >= r1:
                if suf == "ative":
                    if len(w) - len(suf) >= r2:
                        w = w[: -len(suf)]
                else:
                    w = w[: -len(suf)] + rep
            break
    # step 4
    for suf in _STEP4:
        if w.endswith(suf):
            if len(w) - len(suf) >= r2:
                if suf == "ion":
                    if len(w) > 3 and w[-4] in "st":
                        w = w[:-3]
                else:
                    w = w[: -len(suf)]
            break
    # step 5
    if w.endswith("e"):
        if len(w) - 1 >= r2 or (len(w) - 1 >= r1 and not _ends_short_syllable(w[:-1])):
            w = w[:-1]
    elif w.endswith("l"):
        if len(w) - 1 >= r2 and w[:-1].endswith("l"):
            w = w[:-1]
    return w.replace("Y", "y")


# ------------------------------------------------------------------------------ tokenise
def tokenize_py(text: str) -> list[str]:
    """bm25s-style tokens of one text: lower-case, \\w\\w+ words, stop words removed, stemmed.  Pure Python: the
    Unicode-complete restatement and the spec of the native path."""
    return [stem(t) for t in TOKEN_PATTERN.findall(text.lower()) if t not in STOPWORDS_EN]


_lib = None


def _native_lib():
    """libkaito_rag.so for the host-side analysis functions (csrc/text_native.cpp); False when it cannot be loaded -- the
    analysis is host code, so unlike the search path it may run in pure Python."""
    global _lib
    if _lib is None:
        try:
            from . import _native
            _lib = _native.load()
        except Exception:
            _lib = False
    return _lib


def tokenize(text: str) -> list[str]:
    """tokenize_py, through the native analyser (same output, ~20x faster) when the text is ASCII."""
    L = _native_lib()
    if not L or not text.isascii():
        return tokenize_py(text)
    import ctypes as C
    raw = text.encode("ascii")
    cap = len(raw) + 16                       # stems are never longer than their tokens plus one 'e' per token... see below
    n_len, n_tok = C.c_int64(0), C.c_int32(0)
    buf = C.create_string_buffer(cap)
    if L.krag_text_analyze(raw, len(raw), buf, cap, C.byref(n_len), C.byref(n_tok)) != 0:
        return tokenize_py(text)
    if n_len.value > cap:                     # (a token of >= 2 chars yields a stem of at most len + 1 chars and a separator)
        cap = n_len.value
        buf = C.create_string_buffer(cap)
        L.krag_text_analyze(raw, len(raw), buf, cap, C.byref(n_len), C.byref(n_tok))
    return buf.raw[: n_len.value].decode("ascii").split("\n") if n_tok.value else []


class Vocabulary:
    """term string <-> dense u32 id in first-seen order (the host owns the vocabulary)."""

    def __init__(self):
        self.ids: dict[str, int] = {}
        self.terms: list[str] = []

    def __len__(self):
        return len(self.terms)

    def add(self, term: str) -> int:
        i = self.ids.get(term)
        if i is None:
            i = len(self.terms)
            self.ids[term] = i
            self.terms.append(term)
        return i

    def doc_terms(self, text: str):
        """(unique term ids u32, tf u16, doc_len) of one node; grows the vocabulary."""
        toks = tokenize(text)
        cnt = Counter(toks)                     # C-speed counting of the strings; keys keep first-occurrence order, so new terms
        get, add = self.ids.get, self.add       # get their ids in exactly the order a token-by-token pass would hand them out
        known = [get(t) for t in cnt]
        if None in known:
            known = [add(t) if i is None else i for t, i in zip(cnt, known)]
        ids = np.array(known, np.uint32)
        tf = np.minimum(np.fromiter(cnt.values(), np.int64, len(cnt)), 65535).astype(np.uint16)
        return ids, tf, len(toks)

    def query_terms(self, text: str) -> np.ndarray:
        """term ids of a query in token order, duplicates kept, unknown terms dropped (bm25s semantics)."""
        return np.asarray([self.ids[t] for t in tokenize(text) if t in self.ids], np.uint32)


# ------------------------------------------------------------------ WordPiece (BERT / bge tokeniser)
import unicodedata  # noqa: E402


def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    if (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126):
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return ((0x4E00 <= cp <= 0x9FFF) or (0x3400 <= cp <= 0x4DBF) or (0x20000 <= cp <= 0x2A6DF) or (0x2A700 <= cp <= 0x2B73F)
            or (0x2B740 <= cp <= 0x2B81F) or (0x2B820 <= cp <= 0x2CEAF) or (0xF900 <= cp <= 0xFAFF) or (0x2F800 <= cp <= 0x2FA1F))


class WordPieceTokenizer:
    """BertTokenizer (uncased) restated: BasicTokenizer (clean, CJK spacing, lower-case, accent strip,
    punctuation split) + greedy longest-match WordPiece with "##" continuations.  Produces the ids the
    embedder takes: [CLS] tokens [SEP], truncated to max_length.  The reference reaches the same
    tokenizer through sentence-transformers (embedding/huggingface_local_embedding.py:34-53)."""

    def __init__(self, vocab: dict[str, int] | list[str], do_lower_case: bool = True, max_length: int = 512):
        tokens = None
        if not isinstance(vocab, dict):
            tokens = list(vocab)
            vocab = {t: i for i, t in enumerate(tokens)}
        self.vocab, self.lower, self.max_length = vocab, do_lower_case, max_length
        self.unk, self.cls, self.sep = vocab["[UNK]"], vocab["[CLS]"], vocab["[SEP]"]
        # native encoder (csrc/text_native.cpp) for ASCII text; needs the ordered token list and tokens without newlines
        self._nat = None
        L = _native_lib()
        if L and tokens is not None and all("\n" not in t for t in tokens):
            import ctypes as C
            blob = "\n".join(tokens).encode("utf-8")
            h = C.c_void_p()
            if L.krag_wordpiece_create(blob, len(blob), 1 if do_lower_case else 0, C.byref(h)) == 0:
                self._nat = h

    def __del__(self):
        if getattr(self, "_nat", None):
            try:
                _native_lib().krag_wordpiece_destroy(self._nat)
            except Exception:
                pass
            self._nat = None

    @classmethod
    def from_file(cls, path: str, **kw):
        with open(path, encoding="utf-8") as f:
            return cls([line.rstrip("\n") for line in f], **kw)

    def _basic(self, text: str) -> list[str]:
        out = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or (unicodedata.category(ch) in ("Cc", "Cf") and ch not in "\t\n\r"):
                continue
            if _is_cjk(cp):
                out.append(f" {ch} ")
            elif ch in " \t\n\r" or unicodedata.category(ch) == "Zs":
                out.append(" ")
            else:
                out.append(ch)
        toks = []
        for tok in "".join(out).split():
            if self.lower:
                tok = tok.lower()
                tok = "".join(c for c in unicodedata.normalize("NFD", tok) if unicodedata.category(c) != "Mn")
            cur = []
            for ch in tok:
                if _is_punct(ch):
                    if cur:
                        toks.append("".join(cur)); cur = []
                    toks.append(ch)
                else:
                    cur.append(ch)
            if cur:
                toks.append("".join(cur))
        return toks

    def _wordpiece(self, word: str) -> list[int]:
        if len(word) > 100:
            return [self.unk]
        ids, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end] if start == 0 else "##" + word[start:end]
                if sub in self.vocab:
                    cur = self.vocab[sub]
                    break
                end -= 1
            if cur is None:
                return [self.unk]
            ids.append(cur)
            start = end
        return ids

    def encode_py(self, text: str) -> list[int]:
        ids = [i for w in self._basic(text) for i in self._wordpiece(w)]
        return [self.cls] + ids[: self.max_length - 2] + [self.sep]

    def encode(self, text: str) -> list[int]:
        return self.encode_batch([text])[0]

    def encode_batch_flat(self, texts: list[str]):
        """(flat int32 ids, int32 offsets [n + 1]) of every text -- what the embedder's C entry point takes.  All-ASCII batches
        (the common case for queries) never become Python lists: one native call, one masked gather."""
        if self._nat is not None and texts and all(t.isascii() for t in texts):
            import ctypes as C
            raws = [t.encode("ascii") for t in texts]
            offs = np.zeros(len(raws) + 1, np.int64)
            np.cumsum([len(r) for r in raws], out=offs[1:])
            ids = np.empty((len(raws), self.max_length), np.int32)
            cnt = np.empty(len(raws), np.int32)
            rc = _native_lib().krag_wordpiece_encode_batch(self._nat, len(raws), b"".join(raws), offs.ctypes.data_as(C.c_void_p),
                                                           self.max_length, ids.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
            if rc == 0:
                width = int(cnt.max())
                flat = ids[:, :width][np.arange(width, dtype=np.int32)[None, :] < cnt[:, None]]
                toff = np.zeros(len(raws) + 1, np.int32)
                np.cumsum(cnt, out=toff[1:])
                return np.ascontiguousarray(flat, np.int32), toff
        lists = self.encode_batch(texts)
        toff = np.zeros(len(lists) + 1, np.int32)
        np.cumsum([len(t) for t in lists], out=toff[1:])
        flat = np.fromiter((i for t in lists for i in t), np.int32, int(toff[-1]))
        return flat, toff

    def encode_batch(self, texts: list[str]) -> list[list[int]]:
        """ids of every text; ASCII texts go through the native, multi-threaded encoder in one call, the rest through encode_py"""
        out: list = [None] * len(texts)
        idx = [i for i, t in enumerate(texts) if self._nat is not None and t.isascii()]
        if idx:
            import ctypes as C
            raws = [texts[i].encode("ascii") for i in idx]
            offs = np.zeros(len(raws) + 1, np.int64)
            np.cumsum([len(r) for r in raws], out=offs[1:])
            ids = np.empty((len(raws), self.max_length), np.int32)
            cnt = np.empty(len(raws), np.int32)
            rc = _native_lib().krag_wordpiece_encode_batch(self._nat, len(raws), b"".join(raws), offs.ctypes.data_as(C.c_void_p),
                                                           self.max_length, ids.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
            if rc == 0:
                for j, i in enumerate(idx):
                    out[i] = ids[j, : cnt[j]].tolist()
        for i, t in enumerate(texts):
            if out[i] is None:
                out[i] = self.encode_py(t)
        return out
