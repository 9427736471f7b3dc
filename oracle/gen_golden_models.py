"""Generate tests/golden/wire_models_reference.json: the JSON schemas of the reference's own request/response models
(presets/ragengine/models.py, executed unmodified; `llama_index...ChatMessage/MessageRole` -- used only by a chat helper --
are replaced by inert stand-ins, `ragengine.config` is the reference's).  The service's wire models are compared with these
field by field (tests/test_service.py).  Run: python oracle/gen_golden_models.py   (needs /root/reference)."""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "wire_models_reference.json")
MODELS = ["Document", "IndexRequest", "UpdateDocumentRequest", "DeleteDocumentRequest", "RetrieveRequest", "NodeWithScore",
          "RetrieveResponse", "ListDocumentsResponse", "UpdateDocumentResponse", "DeleteDocumentResponse", "HealthStatus"]


def main():
    sys.path.insert(0, "/root/reference/presets")
    for n in ["llama_index", "llama_index.core", "llama_index.core.base", "llama_index.core.base.llms", "llama_index.core.base.llms.types"]:
        sys.modules[n] = types.ModuleType(n)
    t = sys.modules["llama_index.core.base.llms.types"]
    t.ChatMessage = type("ChatMessage", (), {})
    t.MessageRole = type("MessageRole", (), {})
    spec = importlib.util.spec_from_file_location("ref_models", "/root/reference/presets/ragengine/models.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    out = {}
    for name in MODELS:
        cls = getattr(m, name, None)
        if cls is None:
            continue
        sch = cls.model_json_schema()
        out[name] = {"required": sorted(sch.get("required", [])),
                     "properties": {k: {kk: vv for kk, vv in v.items() if kk in ("type", "default", "minimum", "maximum", "anyOf", "items", "$ref")}
                                    for k, v in sch.get("properties", {}).items()}}
    json.dump({"meta": {"source": "presets/ragengine/models.py executed unmodified", "generator": "oracle/gen_golden_models.py"},
               "models": out}, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
