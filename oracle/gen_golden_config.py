"""Generate tests/golden/config_reference.json: the defaults of the reference's environment configuration
(presets/ragengine/config.py executed unmodified with the RAG/LLM/embedding variables unset).
Run: python oracle/gen_golden_config.py   (needs /root/reference)."""
import importlib.util
import json
import os

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "config_reference.json")
KEYS = ["EMBEDDING_SOURCE_TYPE", "LOCAL_EMBEDDING_MODEL_ID", "LLM_INFERENCE_URL", "LLM_ACCESS_SECRET", "LLM_CONTEXT_WINDOW", "VECTOR_DB_TYPE",
        "DEFAULT_VECTOR_DB_PERSIST_DIR", "RAG_SIMILARITY_THRESHOLD", "RAG_DEFAULT_CONTEXT_TOKEN_FILL_RATIO",
        "RAG_DOCUMENT_NODE_TOKEN_APPROXIMATION", "RAG_MAX_TOP_K"]


def main():
    for k in list(os.environ):
        if k.startswith(("RAG_", "LLM_", "EMBEDDING_", "LOCAL_EMBEDDING", "VECTOR_DB", "DEFAULT_VECTOR_DB")):
            del os.environ[k]
    spec = importlib.util.spec_from_file_location("ref_config", "/root/reference/presets/ragengine/config.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    json.dump({"meta": {"source": "presets/ragengine/config.py executed unmodified, variables unset", "generator": "oracle/gen_golden_config.py"},
               "defaults": {k: getattr(m, k) for k in KEYS}}, open(OUT, "w"), indent=1, sort_keys=True)
    print({k: getattr(m, k) for k in KEYS})


if __name__ == "__main__":
    main()
