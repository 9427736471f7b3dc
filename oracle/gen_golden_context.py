"""Generate tests/golden/context_selection_reference.json by RUNNING the reference's own
presets/ragengine/vector_store/node_processors/contex_selection_node_processor.py (ContextSelectionProcessor.__init__ and
_postprocess_nodes) in this container.

llama_index is not installed offline: the six names the module imports from it are replaced by inert stand-ins (plain
containers / no-op pydantic markers); the LLM stand-in only provides `metadata.context_window` and a `count_tokens` that is
the reference's own offline fallback `int(len(text) / 3)` (inference/inference.py:517-521).  The budget arithmetic, the
distance-ascending order, the threshold and the greedy packing recorded in the fixture are executed by the unmodified
reference source read from /root/reference at generation time.  Run:

    python oracle/gen_golden_context.py      # needs /root/reference; writes tests/golden/

Nothing reads /root/reference at test time.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/presets/ragengine/vector_store/node_processors/contex_selection_node_processor.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "context_selection_reference.json")


def _install_stubs():
    class BaseNodePostprocessor:
        def __init__(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

    class NodeWithScore:
        def __init__(self, text, score, nid):
            self.text, self.score, self.nid = text, score, nid

    class QueryBundle:
        def __init__(self, query_str):
            self.query_str = query_str

    names = ["llama_index", "llama_index.core", "llama_index.core.bridge", "llama_index.core.bridge.pydantic", "llama_index.core.llms",
             "llama_index.core.llms.llm", "llama_index.core.postprocessor", "llama_index.core.postprocessor.types",
             "llama_index.core.schema", "llama_index.core.settings"]
    mods = {n: types.ModuleType(n) for n in names}
    mods["llama_index.core.bridge.pydantic"].Field = lambda *a, **k: None
    mods["llama_index.core.bridge.pydantic"].PrivateAttr = lambda *a, **k: None
    mods["llama_index.core.llms.llm"].LLM = object
    mods["llama_index.core.postprocessor.types"].BaseNodePostprocessor = BaseNodePostprocessor
    mods["llama_index.core.schema"].NodeWithScore = NodeWithScore
    mods["llama_index.core.schema"].QueryBundle = QueryBundle
    mods["llama_index.core.settings"].Settings = types.SimpleNamespace(llm=None)
    sys.modules.update(mods)
    return NodeWithScore, QueryBundle


class _LLM:
    def __init__(self, window):
        self.metadata = types.SimpleNamespace(context_window=window)

    def count_tokens(self, text):
        return int(len(text) / 3)


def main():
    NodeWithScore, QueryBundle = _install_stubs()
    spec = importlib.util.spec_from_file_location("ref_context_selection", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = np.random.default_rng(20260921)
    cases = []
    for ci in range(40):
        n = int(g.integers(0, 40))
        window = int(g.choice([200, 1000, 4000, 64000]))
        ratio = float(g.choice([0.2, 0.5, 0.8]))
        max_tokens = None if g.random() < 0.4 else int(g.integers(50, 3000))
        threshold = None if g.random() < 0.2 else float(g.choice([0.85, 0.5, 1.5]))
        query = "q" * int(g.integers(3, 900))
        lens = [int(x) for x in g.integers(1, 1500, n)]
        scores = [float(np.float32(x)) for x in g.random(n) * 2.0]
        if n > 3 and g.random() < 0.5:
            scores[1] = scores[0]                      # equal distances: Python's sort is stable, input order decides
        nodes = [NodeWithScore("t" * L, s, i) for i, (L, s) in enumerate(zip(lens, scores))]
        proc = ref.ContextSelectionProcessor(rag_context_token_fill_ratio=ratio, llm=_LLM(window), max_tokens=max_tokens,
                                             similarity_threshold=threshold)
        picked = proc._postprocess_nodes(nodes, QueryBundle(query))
        cases.append({"window": window, "ratio": ratio, "max_tokens": max_tokens, "threshold": threshold, "query_len": len(query),
                      "text_lens": lens, "scores": scores, "selected": [p.nid for p in picked]})
    meta = {"source": "presets/ragengine/vector_store/node_processors/contex_selection_node_processor.py:31-121 executed unmodified",
            "generator": "oracle/gen_golden_context.py", "addition_prompt_tokens": ref.ADDITION_PROMPT_TOKENS,
            "count_tokens": "int(len(text) / 3) (inference.py:517-521 fallback)"}
    json.dump({"meta": meta, "cases": cases}, open(OUT, "w"), indent=0)
    print(f"wrote {len(cases)} cases to {OUT}; non-empty selections: {sum(1 for c in cases if c['selected'])}")


if __name__ == "__main__":
    main()
