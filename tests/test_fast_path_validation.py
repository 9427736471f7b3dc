"""The HTTP fast paths (kaito_b200/fast_retrieve.py in the single process, kaito_b200/frontend.py in the worker processes) answer a
POST /retrieve themselves only when models.RetrieveRequest (presets/ragengine/models.py; restated in kaito_b200/service.py and
pinned on the reference's schema by tests/test_service.py) would accept the body with the same field values -- everything else
must fall through to FastAPI so the reference's 422 bodies are produced by the real validator."""
import json

import pytest
from hypothesis import given, settings, strategies as st

from kaito_b200 import frontend
from kaito_b200.fast_retrieve import FastRetrieve
from kaito_b200.service import RAG_MAX_TOP_K, RetrieveRequest

scalars = st.one_of(st.none(), st.booleans(), st.integers(-5, 400), st.floats(allow_nan=False, allow_infinity=False, width=32),
                    st.text(max_size=8))
values = st.recursive(scalars, lambda ch: st.one_of(st.lists(ch, max_size=3), st.dictionaries(st.text(max_size=4), ch, max_size=3)), max_leaves=6)
bodies = st.one_of(
    values,
    st.fixed_dictionaries({}, optional={"index_name": st.one_of(st.text(max_size=6), values), "query": st.one_of(st.text(max_size=12), values),
                                       "max_node_count": st.one_of(st.integers(-2, 305), values), "metadata_filter": values,
                                       "context_token_ratio": values, "extra": values}))


@settings(max_examples=600, deadline=None)
@given(bodies)
def test_fast_paths_accept_only_what_the_request_model_accepts(body):
    raw = json.dumps(body).encode()
    fast = FastRetrieve(None, None, None, RAG_MAX_TOP_K, ())._parse(raw)
    assert frontend.parse_retrieve(raw, RAG_MAX_TOP_K) == fast            # one rule, two processes
    if fast is None:
        return
    m = RetrieveRequest.model_validate(body)                              # must not raise
    assert (m.index_name, m.query, m.max_node_count, m.metadata_filter) == fast


@pytest.mark.parametrize("raw", [b"", b"{", b"null", b"[1]", b'"x"', b'{"index_name":"a"}', b'{"index_name":"a","query":"q","max_node_count":1e2}',
                                 b'{"index_name":"a","query":"q","max_node_count":"7"}', b'\xff\xfe'])
def test_malformed_bodies_go_to_fastapi(raw):
    assert frontend.parse_retrieve(raw, RAG_MAX_TOP_K) is None
    assert FastRetrieve(None, None, None, RAG_MAX_TOP_K, ())._parse(raw) is None
