"""Chunking at /index: what `CustomTransformer` (presets/ragengine/vector_store/transformers/custom_transformer.py:23-55)
delegates to LlamaIndex -- `SentenceSplitter()` by default, `CodeSplitter(language=...)` when a document's metadata says
split_type == "code" (a missing language is a ValueError there, :39-41).

Both classes live in llama-index-core (not vendored, not installable offline): their algorithms are restated here from
the published sources [3P-unverified], with the two external pieces they lean on made pluggable:

  token counting   LlamaIndex counts tiktoken tokens (gpt-3.5-turbo encoding).  `default_token_counter()` uses tiktoken when
                   its BPE table can be loaded (it needs a cached file; there is no network in the pod) and otherwise a
                   GPT-2-style pre-tokeniser count (words, numbers, punctuation runs) -- within a few percent on prose.
  sentences        LlamaIndex uses nltk's Punkt tokenizer; `split_sentences` is a regex restatement (terminator + space +
                   capital/quote/digit), keeping the whitespace with the sentence it follows, as span_tokenize does.
  code             CodeSplitter walks a tree-sitter syntax tree.  With `tree_sitter_language_pack` / `tree_sitter_languages`
                   importable the same walk runs on the real tree; otherwise top-level blocks are found from indentation and
                   blank lines (the same max_chars = 1500 packing), which keeps functions/classes together for the common
                   brace- and indent-structured languages.
"""
from __future__ import annotations

import re
from dataclasses import dataclass

_PRETOK = re.compile(r"'s|'t|'re|'ve|'m|'ll|'d| ?[A-Za-z]+| ?\d{1,3}| ?[^\sA-Za-z\d]+|\s+(?!\S)|\s+")


def default_token_counter():
    try:
        import tiktoken
        enc = tiktoken.encoding_for_model("gpt-3.5-turbo")          # llama_index.core.utils.get_tokenizer()
        return lambda text: len(enc.encode(text, allowed_special="all"))
    except Exception:                                                # BPE table not cached: pre-tokeniser pieces, long words count double
        def count(text, _find=_PRETOK.findall):
            ms = _find(text)
            return len(ms) + sum(1 for m in ms if len(m) > 9)
        return count


def _split_keep_separator(text: str, sep: str) -> list[str]:
    parts = text.split(sep)
    return [s for s in ((sep + p) if i > 0 else p for i, p in enumerate(parts)) if s]


_SENT_END = re.compile(r"""(?:(?<=[.!?])|(?<=[.!?]["')\]]))\s+(?=["'(\[]?[A-Z0-9])|(?<=[.!?])\s*\n+\s*""")


def split_sentences(text: str) -> list[str]:
    """sentence spans with their trailing whitespace (nltk PunktSentenceTokenizer.span_tokenize, restated)"""
    out, pos = [], 0
    for m in _SENT_END.finditer(text):
        out.append(text[pos:m.end()])
        pos = m.end()
    if pos < len(text):
        out.append(text[pos:])
    return [s for s in out if s]


@dataclass
class _Split:
    text: str
    is_sentence: bool
    token_size: int


class SentenceSplitter:
    """llama_index.core.node_parser.SentenceSplitter(): chunk_size 1024 tokens, chunk_overlap 200, paragraph separator
    "\\n\\n\\n", then sentences, then the secondary regex, then words, then characters; greedy merge with overlap."""

    CHUNKING_REGEX = "[^,.;。？！]+[,.;。？！]?|[,.;。？！]"

    def __init__(self, chunk_size: int = 1024, chunk_overlap: int = 200, separator: str = " ", paragraph_separator: str = "\n\n\n",
                 token_counter=None):
        if chunk_overlap > chunk_size:
            raise ValueError(f"Got a larger chunk overlap ({chunk_overlap}) than chunk size ({chunk_size}), should be smaller.")
        self.chunk_size, self.chunk_overlap, self.separator, self.paragraph_separator = chunk_size, chunk_overlap, separator, paragraph_separator
        self._count = token_counter or default_token_counter()
        self._split_fns = [lambda t: _split_keep_separator(t, self.paragraph_separator), split_sentences]
        self._sub_sentence_split_fns = [lambda t: re.findall(self.CHUNKING_REGEX, t), lambda t: _split_keep_separator(t, self.separator), list]

    # ---- public
    def split(self, text: str, metadata_str: str = "") -> list[str]:
        """split_text_metadata_aware: the metadata that is prepended to every chunk for the embedder counts against the budget"""
        chunk_size = self.chunk_size
        if metadata_str:
            chunk_size -= self._count(metadata_str)
            if chunk_size <= 0:
                raise ValueError(f"Metadata length is longer than chunk size ({self.chunk_size}). Consider increasing the chunk size or decreasing the size of your metadata.")
        if text == "":
            return [text]
        return self._merge(self._split(text, chunk_size), chunk_size)

    # ---- internals (SentenceSplitter._split / _get_splits_by_fns / _merge)
    def _split(self, text: str, chunk_size: int) -> list[_Split]:
        size = self._count(text)
        if size <= chunk_size:
            return [_Split(text, True, size)]
        pieces, is_sentence = self._splits_by_fns(text)
        out: list[_Split] = []
        for p in pieces:
            n = self._count(p)
            if n <= chunk_size:
                out.append(_Split(p, is_sentence, n))
            else:
                out.extend(self._split(p, chunk_size))
        return out

    def _splits_by_fns(self, text: str):
        for fn in self._split_fns:
            s = fn(text)
            if len(s) > 1:
                return s, True
        for fn in self._sub_sentence_split_fns:
            s = fn(text)
            if len(s) > 1:
                break
        return s, False

    def _merge(self, splits: list[_Split], chunk_size: int) -> list[str]:
        chunks: list[str] = []
        cur: list[tuple[str, int]] = []
        last: list[tuple[str, int]] = []
        cur_len, new_chunk = 0, True

        def close():
            nonlocal cur, last, cur_len, new_chunk
            chunks.append("".join(t for t, _ in cur))
            last, cur, cur_len, new_chunk = cur, [], 0, True
            i = len(last) - 1                                   # overlap: tail of the previous chunk that fits chunk_overlap
            while i >= 0 and cur_len + last[i][1] <= self.chunk_overlap:
                cur.insert(0, last[i]); cur_len += last[i][1]; i -= 1

        splits = list(splits)
        while splits:
            s = splits[0]
            if s.token_size > chunk_size:
                raise ValueError("Single token exceeded chunk size")
            if cur_len + s.token_size > chunk_size and not new_chunk:
                close()
            elif s.is_sentence or cur_len + s.token_size <= chunk_size or new_chunk:
                cur_len += s.token_size
                cur.append((s.text, s.token_size))
                splits.pop(0)
                new_chunk = False
            else:
                close()
        if not new_chunk:
            chunks.append("".join(t for t, _ in cur))
        return [c.strip() for c in chunks if c.strip() != ""]


class CodeSplitter:
    """llama_index.core.node_parser.CodeSplitter(language): chunk_lines 40, chunk_lines_overlap 15, max_chars 1500.  Children of
    the syntax tree are packed into chunks of at most max_chars characters; a child larger than that is split recursively."""

    def __init__(self, language: str, chunk_lines: int = 40, chunk_lines_overlap: int = 15, max_chars: int = 1500):
        self.language, self.chunk_lines, self.chunk_lines_overlap, self.max_chars = language, chunk_lines, chunk_lines_overlap, max_chars
        self._parser = None
        for mod in ("tree_sitter_language_pack", "tree_sitter_languages"):
            try:
                self._parser = __import__(mod).get_parser(language)
                break
            except ImportError:
                continue
            except Exception as e:                               # the library exists but does not know the language
                raise ValueError(f"Could not get parser for language {language}: {e}")

    # the reference's walk over a real tree
    def _chunk_node(self, node, text: bytes, last_end: int = 0) -> list[str]:
        new_chunks, current = [], ""
        for child in node.children:
            if child.end_byte - child.start_byte > self.max_chars:
                if current:
                    new_chunks.append(current); current = ""
                new_chunks.extend(self._chunk_node(child, text, last_end))
            elif len(current) + child.end_byte - child.start_byte > self.max_chars:
                new_chunks.append(current)
                current = text[last_end:child.end_byte].decode("utf-8", "replace")
            else:
                current += text[last_end:child.end_byte].decode("utf-8", "replace")
            last_end = child.end_byte
        if current:
            new_chunks.append(current)
        return new_chunks

    # without tree-sitter: top-level blocks from indentation / blank lines, same packing
    def _blocks(self, text: str) -> list[str]:
        lines = text.splitlines(keepends=True)
        blocks, cur = [], ""
        for i, ln in enumerate(lines):
            top = bool(ln.strip()) and not ln[0].isspace() and not ln.lstrip().startswith(("}", ")", "]", "else", "elif", "except", "finally", "catch"))
            prev_blank = i > 0 and not lines[i - 1].strip()
            if top and cur.strip() and (prev_blank or not lines[i - 1][:1].isspace()):
                blocks.append(cur); cur = ""
            cur += ln
        if cur:
            blocks.append(cur)
        return blocks

    def _pack(self, pieces: list[str]) -> list[str]:
        out, cur = [], ""
        for p in pieces:
            if len(p) > self.max_chars:
                if cur:
                    out.append(cur); cur = ""
                sub = p.splitlines(keepends=True)
                out.extend(self._pack(sub) if len(sub) > 1 else [p[i:i + self.max_chars] for i in range(0, len(p), self.max_chars)])
            elif len(cur) + len(p) > self.max_chars:
                out.append(cur); cur = p
            else:
                cur += p
        if cur:
            out.append(cur)
        return out

    def split(self, text: str) -> list[str]:
        if self._parser is not None:
            tree = self._parser.parse(text.encode("utf-8"))
            if not tree.root_node.children or tree.root_node.children[0].type == "ERROR":
                raise ValueError(f"Could not parse code with language {self.language}.")
            chunks = self._chunk_node(tree.root_node, text.encode("utf-8"))
        else:
            chunks = self._pack(self._blocks(text))
        return [c.strip() for c in chunks if c.strip()]
