"""Python mirror of the reference's four search front-ends (src/server/handlers/search.rs) on top of the GPU
`Dataset`: request decoding, defaults, namespace selection, the `per_page` clamp, `"text"` stripping and the response
shapes. The HTTP server itself (axum, routes) is out of scope; these functions are what the handlers
compute between the decoded request and the JSON body, so the parity tests read like the reference's own would.

Every function returns (http_status, json_body)."""
from __future__ import annotations

from dataclasses import dataclass, field
from urllib.parse import unquote_to_bytes

from .dataset import Dataset, perform_search


@dataclass
class AppState:
    """server_main.rs:16-19: the DatasetManager's name -> Dataset map + default namespace (db/config.rs:91-94)."""
    datasets: dict[str, Dataset] = field(default_factory=dict)
    default_namespace: str = "fugu_db"  # main.rs:120-121


def is_targeting_conversations_or_organizations(filters: list[str]) -> bool:
    """handlers/utils.rs:4-15"""
    for f in filters:
        n = f if f.startswith("/") else "/" + f
        if "/conversation" in n or "/organization" in n:
            return True
    return False


def _strip_text(out: dict) -> None:
    for item in out.get("results", []):
        item.pop("text", None)


def _response_json(resp) -> dict:
    """SearchResponse, src/server/types.rs:146-152 (serde field order)"""
    return {"results": [r.to_json() for r in resp.results], "total": resp.total, "page": resp.page,
            "per_page": resp.per_page, "query": resp.query}


def search_endpoint(state: AppState, payload: dict):
    """POST /search (handlers/search.rs:152-207): default dataset only, defaults page 0 / per_page 20, NO clamp
    (per_page = 0 is a 500: TopDocs::with_limit(0) panics in the reference; here FG_ERR_INVALID)."""
    query = payload["query"]
    filters = payload.get("filters") or []
    page_obj = payload.get("page") or {}
    page = page_obj.get("page") if page_obj.get("page") is not None else 0
    per_page = page_obj.get("per_page") if page_obj.get("per_page") is not None else 20
    ds = state.datasets.get(state.default_namespace)
    if ds is None:
        return 500, {"status": "error", "error": "Default dataset not found"}
    try:
        results = ds.search(query, filters, page, per_page)
    except Exception as e:  # noqa: BLE001 - every failure is a 500 with the message (handlers/search.rs:196-205)
        return 500, {"status": "error", "error": f"Search failed: {e}"}
    return 200, {"status": "success", "query": query, "filters": filters, "page": page, "per_page": per_page,
                 "total": len(results), "results": [r.to_json() for r in results]}


def query_json_post(state: AppState, payload: dict, url_text: bool | None = None, url_include_data: bool | None = None):
    """POST /search/json (handlers/search.rs:210-301): namespace from the body, url flag wins over body flag for
    `text`, developer_message on disagreement, include_data default = not targeting conversations/organizations."""
    body_text = payload.get("text")
    include_text = url_text if url_text is not None else bool(body_text)
    developer_message = None
    if url_text is not None and body_text is not None and url_text != body_text:
        developer_message = "url and request body are set to different values; using url:true/false"
    filters = payload.get("filters") or []
    page_obj = payload.get("page") or {}
    page = page_obj.get("page") if page_obj.get("page") is not None else 0
    per_page = page_obj.get("per_page") if page_obj.get("per_page") is not None else 20
    targeting = is_targeting_conversations_or_organizations(filters)
    include_data = payload.get("include_data")
    if include_data is None:
        include_data = url_include_data
    if include_data is None:
        include_data = not targeting
    namespace = payload.get("namespace") or state.default_namespace
    try:
        resp = perform_search(state.datasets, namespace, payload["query"], filters, page, per_page)
    except Exception as e:  # noqa: BLE001
        return 500, {"error": f"Search failed: {_err(e)}"}
    out = _response_json(resp)
    if not include_text:
        _strip_text(out)
    if developer_message:
        out["developer_message"] = developer_message
    out["includes_data_objects"] = include_data
    out["targeting_conversations_or_organizations"] = targeting
    return 200, out


def query_text_get(state: AppState, q: str, limit: int | None = None, text: bool | None = None, namespace: str | None = None):
    """GET /search?q=&limit=&text=&namespace= (handlers/search.rs:27-76): page 0, per_page = limit (default 20)."""
    try:
        resp = perform_search(state.datasets, namespace or state.default_namespace, q, [], 0, 20 if limit is None else limit)
    except Exception as e:  # noqa: BLE001
        return 500, {"error": f"Search failed: {_err(e)}"}
    out = _response_json(resp)
    if not text:
        _strip_text(out)
    return 200, out


def query_text_path(state: AppState, encoded_query: str, text: bool | None = None, namespace: str | None = None):
    """GET /search/{query} (handlers/search.rs:79-139): URL-decoded query, page 0, per_page 20."""
    try:
        query = unquote_to_bytes(encoded_query).decode("utf-8")
    except UnicodeDecodeError:
        return 400, {"error": "Invalid URL encoding in query"}
    try:
        resp = perform_search(state.datasets, namespace or state.default_namespace, query, [], 0, 20)
    except Exception as e:  # noqa: BLE001
        return 500, {"error": f"Search failed: {_err(e)}"}
    out = _response_json(resp)
    if not text:
        _strip_text(out)
    return 200, out


def _err(e: Exception) -> str:
    # perform_search wraps dataset errors as "Search failed: {e}" (handlers/search.rs:391-397); a missing
    # namespace is reported as is (:365-367)
    if isinstance(e, KeyError):
        return str(e.args[0])
    return f"Search failed: {e}"
